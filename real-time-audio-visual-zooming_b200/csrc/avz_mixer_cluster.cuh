// Cluster-resident far-field mixer: one thread-block cluster of 8 CTAs mixes one utterance with its spectra never
// leaving the chip.  Included by avz_mixer.cu (same translation unit: MixParams, ramp(), kMaxSrc live there).
//
// Reference: rt_av_zoom/core/tf_lite_version/world_building.py:46-52 (apply_frac_delay) and :61-93 (mix_and_save) -
// the same arithmetic as the multi-pass path of avz_mixer.cu (pack two real sources per complex transform, per-bin
// ramps, Hermitian re-pack, two inverse transforms, peak normalisation), with a different transform:
//
//   * the two complex planes of an utterance (sources 0+1 and 2+3 on the way in, mic1 + i mic2 and target + i interferer
//     on the way out) live in the shared memory of 2 x 4 CTAs: CTA (p, a) holds the contiguous quarter
//     x[a M .. (a+1) M) of plane p, M = L / 4 (64 000 samples: 125 KB per CTA) - loads and stores are plain contiguous
//     runs, every sample crosses HBM once in and once out;
//   * the length-L transform is ONE in-place mixed-radix decimation-in-frequency FFT, L = 4 * r1 * r2 * ...: the first
//     radix-4 stage has stride M, i.e. its butterflies take one element from each CTA of the plane - it runs over
//     distributed shared memory (every CTA does the butterflies of a quarter of the positions, reads and writes its
//     three peers in place); all later stages have strides < M and are local to a CTA (radix 16 / 8 / 4 / 2 / 3 / 5 / 7
//     butterflies in registers, one shared-memory round trip per stage);
//   * in-place DIF leaves bin k = d0 + R0 (d1 + R1 (d2 + ...)) at position d0 q0 + d1 q1 + ... (digits reversed); the
//     combine step needs bin k next to bin L - k: a table gives every position its bin and the position of its mirror
//     (quarter u of a plane holds the bins k = u mod 4, so quarters 1 and 3 mirror each other, 0 and 2 themselves);
//     pairs are dealt out evenly over the 8 CTAs, each pair is read and written in place through DSMEM by one thread;
//   * the inverse is the exact reverse (conjugate twiddle on the way in, conjugate butterfly, stages backwards:
//     decimation in time) and lands in natural order; max|mix| is exchanged through DSMEM, the division by it happens
//     on the way out - no scale pass, no atomics.
//
// Index algebra validated in tools/mixer_cluster_model.py before this was written.
#pragma once
#include <cooperative_groups.h>

namespace avz {
namespace {
namespace cgx = cooperative_groups;

constexpr int kClThreads = 512;
constexpr int kClSize = 8;            // 2 planes x 4 quarters
constexpr int kClMaxStages = 12;
constexpr int kClMaxLocal = 25600;    // complex elements per CTA (200 KB)

struct ClStages {
  int L, M, n;
  int radix[kClMaxStages];
  int q[kClMaxStages];        // butterfly stride of the stage (block length radix * q)
  int tw_off[kClMaxStages];   // offset of the stage's twiddles [u-1][pos] in the table (stages with q == 1 have none)
  int pair_off[kClSize + 1];  // CTA r combines pairs [pair_off[r], pair_off[r + 1]) of the pair list
  int n_hi;                   // entries of a ramp table's coarse half: (L / 2 >> 8) + 1
  int tw_sm[kClMaxStages];    // offset of the stage's twiddles in the shared-memory copy, -1: read from global memory
  int tw_sm_n;                // entries of that copy (the small tables of the late stages)
};
constexpr int kTwSmemMax = 2560;   // 20 KB
#ifndef AVZ_CL_CROSSU
#define AVZ_CL_CROSSU 1
#endif
#ifndef AVZ_CL_PAIRU
#define AVZ_CL_PAIRU 1
#endif
#ifndef AVZ_CL_U_SMALL
#define AVZ_CL_U_SMALL 1
#endif
#ifndef AVZ_CL_U_8
#define AVZ_CL_U_8 1
#endif
constexpr int kCrossU = AVZ_CL_CROSSU;    // cross-CTA butterflies a thread has in flight
constexpr int kPairU = AVZ_CL_PAIRU;      // pairs of the combine step a thread has in flight
constexpr int kRampLo = 256;  // exp(-2 pi i k c) = hi[k >> 8] * lo[k & 255]
__host__ __device__ inline size_t cl_ramp_bytes(int n_hi) { return (size_t)2 * 4 * (kRampLo + n_hi) * sizeof(float2); }

// Cluster barrier with release / acquire at cluster scope: what the phases need (a CTA's shared-memory writes visible
// to its peers afterwards).  cooperative_groups' cluster.sync() adds a MEMBAR.ALL.GPU in front, 3-4 % of this kernel.
__device__ __forceinline__ void cl_barrier() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// x / d for a loop-invariant d: r = RN(1 / d) once, then q = RN(x r), q' = RN(q + RN(x - q d) r) - Markstein's
// correction step: the correctly rounded quotient (what __fdiv_rn gives) for three multiply-adds instead of a divide.
__device__ __forceinline__ float div_by(float x, float d, float r) {
  const float q = x * r;
  return fmaf(fmaf(-q, d, x), r, q);
}
template <bool INV>
__device__ __forceinline__ float2 mul_mi(float2 a) {   // * (-i) forward, * (+i) inverse
  return INV ? make_float2(-a.y, a.x) : make_float2(a.y, -a.x);
}
template <bool INV>
__device__ __forceinline__ float2 mul_w(float2 a, float c, float s) {   // * (c - i s) forward, * (c + i s) inverse
  return INV ? cmulc(a, make_float2(c, -s)) : cmul(a, make_float2(c, -s));
}
template <bool INV>
__device__ __forceinline__ void bfly4(float2& v0, float2& v1, float2& v2, float2& v3) {
  const float2 t0 = cadd(v0, v2), t1 = csub(v0, v2), t2 = cadd(v1, v3), t3 = mul_mi<INV>(csub(v1, v3));
  v0 = cadd(t0, t2);
  v1 = cadd(t1, t3);
  v2 = csub(t0, t2);
  v3 = csub(t1, t3);
}
__device__ __forceinline__ void bfly2(float2& v0, float2& v1) {
  const float2 a = cadd(v0, v1), b = csub(v0, v1);
  v0 = a;
  v1 = b;
}

template <int R>
__device__ __forceinline__ float odd_cos(int j) {   // cos(2 pi j / R), j <= (R-1)/2 (compile-time j after unrolling)
  if (R == 3) return j == 0 ? 1.f : -0.5f;
  if (R == 5) return j == 0 ? 1.f : (j == 1 ? 0.30901699437494745f : -0.8090169943749475f);
  return j == 0 ? 1.f : (j == 1 ? 0.6234898018587336f : (j == 2 ? -0.2225209339563144f : -0.9009688679024191f));
}
template <int R>
__device__ __forceinline__ float odd_sin(int j) {
  if (R == 3) return j == 0 ? 0.f : 0.8660254037844386f;
  if (R == 5) return j == 0 ? 0.f : (j == 1 ? 0.9510565162951535f : 0.5877852522924731f);
  return j == 0 ? 0.f : (j == 1 ? 0.7818314824680298f : (j == 2 ? 0.9749279121818236f : 0.4338837391175581f));
}

// Length-R DFT of v in place, natural order in and out (sign - forward, + inverse; unnormalised).
template <int R, bool INV>
__device__ __forceinline__ void bfly(float2 (&v)[R]) {
  if constexpr (R == 2) {
    bfly2(v[0], v[1]);
  } else if constexpr (R == 4) {
    bfly4<INV>(v[0], v[1], v[2], v[3]);
  } else if constexpr (R == 8) {
    // t = 2 t1 + t2, u = u1 + 4 u2: radix 4 over t1, twiddle W8^{t2 u1}, radix 2 over t2
    bfly4<INV>(v[0], v[2], v[4], v[6]);
    bfly4<INV>(v[1], v[3], v[5], v[7]);   // y[t2 = 1][u1] at v[2 u1 + 1]
    constexpr float h = 0.70710678118654752f;
    v[3] = mul_w<INV>(v[3], h, h);
    v[5] = mul_mi<INV>(v[5]);
    v[7] = mul_w<INV>(v[7], -h, h);
    bfly2(v[0], v[1]);
    bfly2(v[2], v[3]);
    bfly2(v[4], v[5]);
    bfly2(v[6], v[7]);   // X[u1] at v[2 u1], X[u1 + 4] at v[2 u1 + 1]
    const float2 x1 = v[2], x2 = v[4], x3 = v[6], x4 = v[1], x5 = v[3], x6 = v[5];
    v[1] = x1; v[2] = x2; v[3] = x3; v[4] = x4; v[5] = x5; v[6] = x6;
  } else if constexpr (R == 16) {
    // t = 4 t1 + t2, u = u1 + 4 u2: radix 4 over t1 (y[t2][u1] at v[4 u1 + t2]), twiddle W16^{t2 u1}, radix 4 over t2
#pragma unroll
    for (int t2 = 0; t2 < 4; ++t2) bfly4<INV>(v[t2], v[4 + t2], v[8 + t2], v[12 + t2]);
    constexpr float c1 = 0.92387953251128674f, s1 = 0.38268343236508977f, h = 0.70710678118654752f;
    v[4 + 1] = mul_w<INV>(v[4 + 1], c1, s1);     // W16^1
    v[4 + 2] = mul_w<INV>(v[4 + 2], h, h);       // W16^2
    v[4 + 3] = mul_w<INV>(v[4 + 3], s1, c1);     // W16^3
    v[8 + 1] = mul_w<INV>(v[8 + 1], h, h);       // W16^2
    v[8 + 2] = mul_mi<INV>(v[8 + 2]);            // W16^4
    v[8 + 3] = mul_w<INV>(v[8 + 3], -h, h);      // W16^6
    v[12 + 1] = mul_w<INV>(v[12 + 1], s1, c1);   // W16^3
    v[12 + 2] = mul_w<INV>(v[12 + 2], -h, h);    // W16^6
    v[12 + 3] = mul_w<INV>(v[12 + 3], -c1, -s1); // W16^9
#pragma unroll
    for (int u1 = 0; u1 < 4; ++u1) bfly4<INV>(v[4 * u1], v[4 * u1 + 1], v[4 * u1 + 2], v[4 * u1 + 3]);
    // X[u1 + 4 u2] sits at v[4 u1 + u2]: transpose the 4 x 4 index grid
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = i + 1; j < 4; ++j) {
        const float2 t = v[4 * i + j];
        v[4 * i + j] = v[4 * j + i];
        v[4 * j + i] = t;
      }
  } else {
    // odd prime: pair x_j with x_{R-j}; y_u = a_u -+ i b_u, y_{R-u} = a_u +- i b_u
    constexpr int H = (R - 1) / 2;
    float2 tp[H + 1], tm[H + 1];
#pragma unroll
    for (int j = 1; j <= H; ++j) {
      tp[j] = cadd(v[j], v[R - j]);
      tm[j] = csub(v[j], v[R - j]);
    }
    const float2 x0 = v[0];
    float2 y0 = x0;
#pragma unroll
    for (int j = 1; j <= H; ++j) y0 = cadd(y0, tp[j]);
    v[0] = y0;
#pragma unroll
    for (int u = 1; u <= H; ++u) {
      float2 a = x0, b = make_float2(0.f, 0.f);
#pragma unroll
      for (int j = 1; j <= H; ++j) {
        const int r = (j * u) % R;
        const float c = odd_cos<R>(r <= H ? r : R - r);
        const float s = (r <= H) ? odd_sin<R>(r) : -odd_sin<R>(R - r);
        a = cfma_real(tp[j], c, a);
        b = cfma_real(tm[j], s, b);
      }
      const float2 ib = mul_mi<INV>(b);
      v[u] = cadd(a, ib);
      v[R - u] = csub(a, ib);
    }
  }
}

// One stage on a CTA's own M elements: butterflies over x[blk R q + pos + q t], t < R.  A thread works on U butterflies
// at a time - all their loads are issued before the first store (the transform is in place, and the compiler cannot
// know that two butterflies never share an element): with four warps per scheduler it is memory-level parallelism
// inside a thread, not occupancy, that hides the shared-memory and twiddle latencies.
template <int R, bool INV, int U>
__device__ __forceinline__ void cl_local_stage(float2* __restrict__ sm, int M, int q, const float2* __restrict__ tw) {
  const int nbf = M / R;
  for (int i0 = threadIdx.x; i0 < nbf; i0 += U * kClThreads) {
    float2 v[U][R];
    float2* x[U];
    int pos[U];
#pragma unroll
    for (int b = 0; b < U; ++b) {
      int idx = i0 + b * kClThreads;
      idx = idx < nbf ? idx : i0;          // a batch's missing tail repeats its first butterfly (loads only)
      int blk = idx;
      pos[b] = 0;
      if (q > 1) {
        blk = idx / q;
        pos[b] = idx - blk * q;
      }
      x[b] = sm + blk * (R * q) + pos[b];
#pragma unroll
      for (int t = 0; t < R; ++t) v[b][t] = x[b][t * q];
    }
#pragma unroll
    for (int b = 0; b < U; ++b) {
      float2 w[R];
      if (q > 1) {
        if constexpr (R == 16) {
          // W^{pos (4 u2 + u1)} = W^{4 pos u2} W^{pos u1}: six loads and nine products instead of fifteen loads
          float2 w1[4], w4[4];
#pragma unroll
          for (int j = 1; j < 4; ++j) {
            w1[j] = tw[(j - 1) * q + pos[b]];
            w4[j] = tw[(4 * j - 1) * q + pos[b]];
          }
#pragma unroll
          for (int u = 1; u < 16; ++u) {
            const int u1 = u & 3, u2 = u >> 2;
            w[u] = (u2 == 0) ? w1[u1] : (u1 == 0 ? w4[u2] : cmul(w4[u2], w1[u1]));
          }
        } else {
#pragma unroll
          for (int u = 1; u < R; ++u) w[u] = tw[(u - 1) * q + pos[b]];
        }
      }
      if (INV && q > 1) {
#pragma unroll
        for (int t = 1; t < R; ++t) v[b][t] = cmulc(v[b][t], w[t]);
      }
      bfly<R, INV>(v[b]);
      if (!INV && q > 1) {
#pragma unroll
        for (int u = 1; u < R; ++u) v[b][u] = cmul(v[b][u], w[u]);
      }
    }
#pragma unroll
    for (int b = 0; b < U; ++b) {
      if (i0 + b * kClThreads < nbf) {
#pragma unroll
        for (int u = 0; u < R; ++u) x[b][u * q] = v[b][u];
      }
    }
  }
}

template <bool INV>
__device__ __forceinline__ void cl_local_dispatch(float2* sm, const ClStages& pl, int s, const float2* __restrict__ t) {
  switch (pl.radix[s]) {
    case 16: cl_local_stage<16, INV, 1>(sm, pl.M, pl.q[s], t); break;
    case 8: cl_local_stage<8, INV, AVZ_CL_U_8>(sm, pl.M, pl.q[s], t); break;
    case 4: cl_local_stage<4, INV, AVZ_CL_U_SMALL>(sm, pl.M, pl.q[s], t); break;
    case 2: cl_local_stage<2, INV, AVZ_CL_U_SMALL>(sm, pl.M, pl.q[s], t); break;
    case 3: cl_local_stage<3, INV, AVZ_CL_U_SMALL>(sm, pl.M, pl.q[s], t); break;
    case 5: cl_local_stage<5, INV, AVZ_CL_U_SMALL>(sm, pl.M, pl.q[s], t); break;
    default: cl_local_stage<7, INV, AVZ_CL_U_8>(sm, pl.M, pl.q[s], t); break;
  }
}

__global__ void __launch_bounds__(kClThreads, 1)
k_mix_cluster(const float* __restrict__ src, int B, int S, ClStages pl, const float2* __restrict__ tw,
              const int4* __restrict__ pairs, MixParams prm, float peak_eps, int vec, int dbg, float* __restrict__ mix,
              float* __restrict__ tgt, float* __restrict__ itf) {
  // dbg (AVZ_EXPERIMENT builds only, 0 otherwise): stop after phase `dbg` and store the planes as they are
  // (1 load, 2 forward stage 0, 3 forward local stages, 4 combine, 5 inverse local stages)
  extern __shared__ __align__(16) unsigned char cl_smem[];
  float2* sm = reinterpret_cast<float2*>(cl_smem);   // this CTA's quarter of its plane: M complex values
  float2* s_ramp = sm + pl.M;                        // [2 (mic)][4 (source)][kRampLo + n_hi] phase-ramp factors
  float2* s_tw = s_ramp + 2 * 4 * (kRampLo + pl.n_hi);   // twiddles of the late local stages
  __shared__ float s_red[kClThreads / 32];
  __shared__ float s_peak;
  cgx::cluster_group cl = cgx::this_cluster();
  const int rank = (int)cl.block_rank(), p = rank >> 2, a = rank & 3;
  const int n_clusters = gridDim.x / kClSize, cid = blockIdx.x / kClSize;
  const int L = pl.L, M = pl.M;
  const int P = (S + 1) / 2;
  const bool active = 2 * p < S;            // S <= 2: plane 1 has nothing to transform on the way in
  const float inv_n = (float)(1.0 / (double)L);
  float2* peer[4];
#pragma unroll
  for (int u = 0; u < 4; ++u) peer[u] = cl.map_shared_rank(sm, (p << 2) + u);
  const float2* tw0 = tw + pl.tw_off[0];    // [3][M]: W_L^{m u}
  const int per = (M + 3) / 4, mlo = a * per, mhi = min(M, mlo + per);   // this CTA's share of the cross-CTA butterflies
  // Phase ramps exp(-2 pi i k c), c = tau fs / L per source and microphone (the delays are per call, not per utterance):
  // k = 256 kh + kl, one table per factor, from a float64 phase reduced to [0, 2) - a ramp then costs two shared-memory
  // loads and one complex multiply instead of a sincospif, at the same ~1e-7 accuracy.
  {
    const int tl = kRampLo + pl.n_hi;
    for (int idx = threadIdx.x; idx < 2 * S * tl; idx += kClThreads) {
      const int t = idx / tl, e = idx - t * tl;          // t = mic * S + source
      const int mic = t / S, sc = t - mic * S;
      const double c = mic ? prm.c2[sc] : prm.c1[sc];
      const double k = (e < kRampLo) ? (double)e : (double)(e - kRampLo) * (double)kRampLo;
      double ph = 2.0 * k * c;
      ph -= 2.0 * floor(0.5 * ph);
      double sn, cs;
      sincospi(ph, &sn, &cs);
      s_ramp[(mic * 4 + sc) * tl + e] = make_float2((float)cs, (float)-sn);
    }
    for (int s = 1; s < pl.n; ++s)
      if (pl.tw_sm[s] >= 0)
        for (int i = threadIdx.x; i < (pl.radix[s] - 1) * pl.q[s]; i += kClThreads)
          s_tw[pl.tw_sm[s] + i] = __ldg(tw + pl.tw_off[s] + i);
    __syncthreads();
  }
  auto stage_tw = [&](int s) -> const float2* { return pl.tw_sm[s] >= 0 ? s_tw + pl.tw_sm[s] : tw + pl.tw_off[s]; };

  for (int b = cid; b < B; b += n_clusters) {
    // ---- load: two real sources -> re + i im, this CTA's quarter
    if (active) {
      const float* sa = src + ((int64_t)b * S + 2 * p) * L + (int64_t)a * M;
      const bool has_b = 2 * p + 1 < S;
      const float* sb = sa + L;
      if (vec) {
        float4* s4 = reinterpret_cast<float4*>(sm);
        for (int i = threadIdx.x; i < M / 4; i += kClThreads) {
          const float4 xa = __ldcs(reinterpret_cast<const float4*>(sa) + i);
          const float4 xb = has_b ? __ldcs(reinterpret_cast<const float4*>(sb) + i) : make_float4(0.f, 0.f, 0.f, 0.f);
          s4[2 * i] = make_float4(xa.x, xb.x, xa.y, xb.y);
          s4[2 * i + 1] = make_float4(xa.z, xb.z, xa.w, xb.w);
        }
      } else {
        for (int i = threadIdx.x; i < M; i += kClThreads) sm[i] = make_float2(__ldcs(sa + i), has_b ? __ldcs(sb + i) : 0.f);
      }
    }
    cl_barrier();
    // ---- forward stage 0 (radix 4, stride M): across the four CTAs of the plane, in place
    if (active && (dbg == 0 || dbg >= 2)) {
      for (int m0 = mlo + threadIdx.x; m0 < mhi; m0 += kCrossU * kClThreads) {
        float2 v[kCrossU][4], w[kCrossU][4];
#pragma unroll
        for (int c = 0; c < kCrossU; ++c) {
          const int m = min(m0 + c * kClThreads, mhi - 1);
#pragma unroll
          for (int u = 0; u < 4; ++u) v[c][u] = peer[u][m];
#pragma unroll
          for (int u = 1; u < 4; ++u) w[c][u] = __ldg(tw0 + (u - 1) * M + m);
        }
#pragma unroll
        for (int c = 0; c < kCrossU; ++c) {
          bfly4<false>(v[c][0], v[c][1], v[c][2], v[c][3]);
#pragma unroll
          for (int u = 1; u < 4; ++u) v[c][u] = cmul(v[c][u], w[c][u]);
        }
#pragma unroll
        for (int c = 0; c < kCrossU; ++c) {
          const int m = m0 + c * kClThreads;
          if (m < mhi) {
#pragma unroll
            for (int u = 0; u < 4; ++u) peer[u][m] = v[c][u];
          }
        }
      }
    }
    // the next utterance's sources: ask L2 for them now, the load at the top of the loop then does not wait for HBM
    if (active && b + n_clusters < B) {
      const char* nx = reinterpret_cast<const char*>(src + ((int64_t)(b + n_clusters) * S + 2 * p) * L + (int64_t)a * M);
      const int n_src = (2 * p + 1 < S) ? 2 : 1;
      const int lines = (int)(((size_t)M * sizeof(float) + 127) / 128);
      for (int i = threadIdx.x; i < n_src * lines; i += kClThreads) {
        const int which = i / lines, ln = i - which * lines;
        asm volatile("prefetch.global.L2 [%0];" ::"l"(nx + (size_t)which * L * sizeof(float) + (size_t)ln * 128));
      }
    }
    cl_barrier();
    // ---- forward local stages
    if (active && (dbg == 0 || dbg >= 3)) {
      for (int s = 1; s < pl.n; ++s) {
        cl_local_dispatch<false>(sm, pl, s, stage_tw(s));
        __syncthreads();
      }
    }
    cl_barrier();
    // ---- combine: unpack the sources at bins (k, L-k), ramps, sums, re-pack - in place on both planes
    // (the pair list gives every CTA an equal share of the L/2 + 1 pairs, sorted by the local position of the smaller
    // bin, which lies in this CTA's quarter - of this plane or of the other one)
    float2* own0 = (p == 0) ? sm : cl.map_shared_rank(sm, a);
    float2* own1 = (p == 1) ? sm : cl.map_shared_rank(sm, 4 + a);
    const int rtl = kRampLo + pl.n_hi;
    const int pr_end = (dbg == 0 || dbg >= 4) ? pl.pair_off[rank + 1] : 0;
    for (int i0 = pl.pair_off[rank] + threadIdx.x; i0 < pr_end; i0 += kPairU * kClThreads) {
      int4 e[kPairU];
      float2 zk[kPairU][2], zm[kPairU][2];
#pragma unroll
      for (int c = 0; c < kPairU; ++c) e[c] = __ldg(pairs + min(i0 + c * kClThreads, pr_end - 1));
#pragma unroll
      for (int c = 0; c < kPairU; ++c) {   // all loads of the batch before its first store (in place, see cl_local_stage)
        const int jo = e[c].x, qm = e[c].y >> 16, jm = e[c].y & 0xffff;
        zk[c][0] = own0[jo];
        zm[c][0] = cl.map_shared_rank(sm, qm)[jm];
        if (P > 1) {
          zk[c][1] = own1[jo];
          zm[c][1] = cl.map_shared_rank(sm, 4 + qm)[jm];
        }
      }
#pragma unroll
      for (int c = 0; c < kPairU; ++c) {
        if (i0 + c * kClThreads >= pr_end) continue;
        const int jo = e[c].x, qm = e[c].y >> 16, jm = e[c].y & 0xffff, kk = e[c].z;
        const bool self = e[c].w != 0;
        float2 m1 = make_float2(0.f, 0.f), m2 = m1, tg = m1;
#pragma unroll
        for (int pp = 0; pp < 2; ++pp) {
          if (pp < P) {
            const float2 av = make_float2(0.5f * (zk[c][pp].x + zm[c][pp].x), 0.5f * (zk[c][pp].y - zm[c][pp].y));
            const float2 bv = make_float2(0.5f * (zk[c][pp].y + zm[c][pp].y), 0.5f * (zm[c][pp].x - zk[c][pp].x));
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              const int s = 2 * pp + h;
              if (s < S) {
                const float2 v = h ? bv : av;
                const float2* t1 = s_ramp + s * rtl;
                const float2 r1 = cmul(t1[kRampLo + (kk >> 8)], t1[kk & (kRampLo - 1)]);
                float2 r2 = make_float2(r1.x, -r1.y);
                if (!prm.sym[s]) {
                  const float2* t2 = s_ramp + (4 + s) * rtl;
                  r2 = cmul(t2[kRampLo + (kk >> 8)], t2[kk & (kRampLo - 1)]);
                }
                const float2 d1 = cmul(v, r1);
                const float2 d2 = cmul(v, r2);
                m1 = cadd(m1, d1);
                m2 = cadd(m2, d2);
                if (s == 0) tg = d1;
              }
            }
          }
        }
        float2 in = csub(m1, tg);
        if (self) {  // DC and Nyquist: the real inverse transform ignores the imaginary part
          m1.y = 0.f; m2.y = 0.f; tg.y = 0.f; in.y = 0.f;
        }
        own0[jo] = make_float2(m1.x - m2.y, m1.y + m2.x);
        own1[jo] = make_float2(tg.x - in.y, tg.y + in.x);
        if (!self) {
          cl.map_shared_rank(sm, qm)[jm] = make_float2(m1.x + m2.y, m2.x - m1.y);
          cl.map_shared_rank(sm, 4 + qm)[jm] = make_float2(tg.x + in.y, in.x - tg.y);
        }
      }
    }
    cl_barrier();
    // ---- inverse local stages, backwards
    for (int s = pl.n - 1; s >= ((dbg == 0 || dbg >= 5) ? 1 : pl.n); --s) {
      cl_local_dispatch<true>(sm, pl, s, stage_tw(s));
      __syncthreads();
    }
    cl_barrier();
    // ---- inverse stage 0 across the CTAs, 1/L, max|mix|
    float mx = 0.f;
    for (int m0 = mlo + threadIdx.x; m0 < (dbg == 0 ? mhi : 0); m0 += kCrossU * kClThreads) {
      float2 v[kCrossU][4], w[kCrossU][4];
#pragma unroll
      for (int c = 0; c < kCrossU; ++c) {
        const int m = min(m0 + c * kClThreads, mhi - 1);
#pragma unroll
        for (int u = 0; u < 4; ++u) v[c][u] = peer[u][m];
#pragma unroll
        for (int u = 1; u < 4; ++u) w[c][u] = __ldg(tw0 + (u - 1) * M + m);
      }
#pragma unroll
      for (int c = 0; c < kCrossU; ++c) {
#pragma unroll
        for (int u = 1; u < 4; ++u) v[c][u] = cmulc(v[c][u], w[c][u]);
        bfly4<true>(v[c][0], v[c][1], v[c][2], v[c][3]);
      }
#pragma unroll
      for (int c = 0; c < kCrossU; ++c) {
        const int m = m0 + c * kClThreads;
        if (m < mhi) {
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const float2 o = make_float2(v[c][u].x * inv_n, v[c][u].y * inv_n);
            mx = fmaxf(mx, fmaxf(fabsf(o.x), fabsf(o.y)));
            peer[u][m] = o;
          }
        }
      }
    }
    mx = warp_max(mx);
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = mx;
    __syncthreads();
    if (threadIdx.x == 0) {
      for (int w = 1; w < kClThreads / 32; ++w) mx = fmaxf(mx, s_red[w]);
      s_peak = mx;
    }
    cl_barrier();
    // ---- store: / (max|mix| + eps) (world_building.py:86-91), contiguous quarter of each of the two signals
    float den = 1.f;
    const bool norm = peak_eps >= 0.f;
    if (norm) {
      float pk = 0.f;
#pragma unroll
      for (int u = 0; u < 4; ++u) pk = fmaxf(pk, *cl.map_shared_rank(&s_peak, u));
      den = pk + peak_eps;
    }
    const float rden = __frcp_rn(den);
    float* o_re = (p == 0 ? mix + (int64_t)b * 2 * L : tgt + (int64_t)b * L) + (int64_t)a * M;
    float* o_im = (p == 0 ? mix + ((int64_t)b * 2 + 1) * L : itf + (int64_t)b * L) + (int64_t)a * M;
    if (vec) {
      const float4* s4 = reinterpret_cast<const float4*>(sm);
      for (int i = threadIdx.x; i < M / 4; i += kClThreads) {
        const float4 lo = s4[2 * i], hi = s4[2 * i + 1];
        float4 re = make_float4(lo.x, lo.z, hi.x, hi.z), im = make_float4(lo.y, lo.w, hi.y, hi.w);
        if (norm) {
          re = make_float4(div_by(re.x, den, rden), div_by(re.y, den, rden), div_by(re.z, den, rden), div_by(re.w, den, rden));
          im = make_float4(div_by(im.x, den, rden), div_by(im.y, den, rden), div_by(im.z, den, rden), div_by(im.w, den, rden));
        }
        __stcs(reinterpret_cast<float4*>(o_re) + i, re);
        __stcs(reinterpret_cast<float4*>(o_im) + i, im);
      }
    } else {
      for (int i = threadIdx.x; i < M; i += kClThreads) {
        const float2 v = sm[i];
        o_re[i] = norm ? div_by(v.x, den, rden) : v.x;
        o_im[i] = norm ? div_by(v.y, den, rden) : v.y;
      }
    }
    __syncthreads();   // everyone has read its quarter before the next utterance's load overwrites it
  }
  cl_barrier();   // nobody leaves while a peer may still be reading its s_peak
}

// ---- host side: plan (radices, twiddle tables, bin / mirror-position table) per signal length and device ----
struct ClusterPlan {
  int device;
  int64_t L;
  ClStages st;
  const float2* tw;
  const int4* pairs;
  int max_clusters;
  size_t smem;
};
std::mutex g_cl_mu;
std::vector<ClusterPlan*> g_cl_plans;

bool cluster_radices(int64_t L, ClStages* st) {
  if (L < 8 || (L & 3) != 0 || L / 4 > kClMaxLocal) return false;
  st->n_hi = (int)((L / 2) >> 8) + 1;
  if ((size_t)(L / 4) * sizeof(float2) + cl_ramp_bytes(st->n_hi) + kTwSmemMax * sizeof(float2) > (size_t)225 * 1024) return false;
  int m = (int)(L / 4), n = 0;
  st->L = (int)L;
  st->M = m;
  st->radix[n++] = 4;
  while (m % 16 == 0 && n < kClMaxStages) { st->radix[n++] = 16; m /= 16; }
  for (int r : {8, 4, 2})
    if (m % r == 0 && n < kClMaxStages) { st->radix[n++] = r; m /= r; }
  for (int r : {3, 5, 7})
    while (m % r == 0 && n < kClMaxStages) { st->radix[n++] = r; m /= r; }
  if (m != 1) return false;
  st->n = n;
  int len = (int)L;
  for (int s = 0; s < n; ++s) {
    len /= st->radix[s];
    st->q[s] = len;
  }
  return true;
}

// nullptr in *out (and AVZ_OK) when this length has no cluster plan or the device cannot run the cluster
int cluster_plan_for(int64_t L, const ClusterPlan** out) {
  *out = nullptr;
  ClStages st{};
  if (!cluster_radices(L, &st)) return AVZ_OK;
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return set_error(AVZ_ENOGPU, "cudaGetDevice: %s", cudaGetErrorString(e));
  std::lock_guard<std::mutex> lk(g_cl_mu);
  for (const ClusterPlan* p : g_cl_plans)
    if (p->device == dev && p->L == L) {
      *out = p->max_clusters > 0 ? p : nullptr;
      return AVZ_OK;
    }
  const double two_pi = 6.283185307179586476925286766559;
  std::vector<float2> tw;
  for (int s = 0; s < st.n; ++s) {
    st.tw_off[s] = (int)tw.size();
    const int R = st.radix[s], q = st.q[s];
    if (q == 1) continue;
    const double nb = (double)R * (double)q;
    for (int u = 1; u < R; ++u)
      for (int pos = 0; pos < q; ++pos) {
        const double ang = two_pi * (double)(((int64_t)pos * u) % ((int64_t)R * q)) / nb;
        tw.push_back(make_float2((float)cos(ang), (float)-sin(ang)));
      }
  }
  if (tw.empty()) tw.push_back(make_float2(1.f, 0.f));
  // the late stages' tables are small and read by every butterfly of every utterance: they get a copy in shared memory
  st.tw_sm_n = 0;
  for (int s = 0; s < kClMaxStages; ++s) st.tw_sm[s] = -1;
  for (int s = st.n - 1; s >= 1; --s) {
    const int n = (st.radix[s] - 1) * st.q[s];
    if (st.q[s] == 1) continue;
    if (st.tw_sm_n + n > kTwSmemMax) break;
    st.tw_sm[s] = st.tw_sm_n;
    st.tw_sm_n += n;
  }
  // bin of every position (digits reversed) and the position of its mirror bin, packed as (quarter << 16 | local index)
  const int Li = (int)L, M = st.M;
  std::vector<int> bin_of((size_t)Li), pos_of((size_t)Li);
  for (int pos = 0; pos < Li; ++pos) {
    int k = 0, mult = 1;
    for (int s = 0; s < st.n; ++s) {
      k += ((pos / st.q[s]) % st.radix[s]) * mult;
      mult *= st.radix[s];
    }
    bin_of[(size_t)pos] = k;
    pos_of[(size_t)k] = pos;
  }
  // pairs (k, L - k), k <= L / 2, grouped by the quarter that holds bin k, sorted by its local position there; each
  // quarter's list is halved between its plane-0 and its plane-1 CTA
  std::vector<int4> pairs;
  {
    std::vector<std::vector<int4>> by_q(4);
    for (int pos = 0; pos < Li; ++pos) {
      const int k = bin_of[(size_t)pos];
      const int km = (Li - k) % Li;
      if (k > km) continue;
      const int mp = pos_of[(size_t)km];
      by_q[(size_t)(pos / M)].push_back(make_int4(pos % M, ((mp / M) << 16) | (mp % M), k, mp == pos ? 1 : 0));
    }
    std::vector<std::vector<int4>> by_rank(kClSize);
    for (int a = 0; a < 4; ++a) {
      const size_t half = (by_q[(size_t)a].size() + 1) / 2;
      for (size_t i = 0; i < by_q[(size_t)a].size(); ++i) by_rank[(size_t)(i < half ? a : 4 + a)].push_back(by_q[(size_t)a][i]);
    }
    for (int r = 0; r < kClSize; ++r) {
      st.pair_off[r] = (int)pairs.size();
      pairs.insert(pairs.end(), by_rank[(size_t)r].begin(), by_rank[(size_t)r].end());
    }
    st.pair_off[kClSize] = (int)pairs.size();
  }
  void *dtw = nullptr, *dkp = nullptr;
  AVZ_CUDA_OK(cudaMalloc(&dtw, tw.size() * sizeof(float2)));
  AVZ_CUDA_OK(cudaMalloc(&dkp, pairs.size() * sizeof(int4)));
  AVZ_CUDA_OK(cudaMemcpy(dtw, tw.data(), tw.size() * sizeof(float2), cudaMemcpyHostToDevice));
  AVZ_CUDA_OK(cudaMemcpy(dkp, pairs.data(), pairs.size() * sizeof(int4), cudaMemcpyHostToDevice));
  ClusterPlan* p = new ClusterPlan{dev, L, st, (const float2*)dtw, (const int4*)dkp, 0,
                                   (size_t)M * sizeof(float2) + cl_ramp_bytes(st.n_hi) + (size_t)st.tw_sm_n * sizeof(float2)};
  // how many such clusters the device runs at once (0: it cannot - the caller falls back to the multi-pass path)
  if (cudaFuncSetAttribute(k_mix_cluster, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p->smem) == cudaSuccess) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(kClSize);
    cfg.blockDim = dim3(kClThreads);
    cfg.dynamicSmemBytes = p->smem;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = kClSize;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    int nc = 0;
    if (cudaOccupancyMaxActiveClusters(&nc, k_mix_cluster, &cfg) == cudaSuccess) p->max_clusters = nc;
  }
  (void)cudaGetLastError();
  g_cl_plans.push_back(p);
  *out = p->max_clusters > 0 ? p : nullptr;
  return AVZ_OK;
}

int launch_mix_cluster(const ClusterPlan* cp, const float* src, int B, int S, const MixParams& prm, float peak_eps,
                       float* mix, float* tgt, float* itf, cudaStream_t st) {
  {
    static std::mutex mu;
    std::lock_guard<std::mutex> lk(mu);
    AVZ_CUDA_OK(cudaFuncSetAttribute(k_mix_cluster, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cp->smem));
  }
  const int n_clusters = B < cp->max_clusters ? B : cp->max_clusters;
  const bool aligned = (((uintptr_t)src | (uintptr_t)mix | (uintptr_t)tgt | (uintptr_t)itf) & 15u) == 0;
  const int vec = (aligned && (cp->st.M & 3) == 0) ? 1 : 0;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(n_clusters * kClSize));
  cfg.blockDim = dim3(kClThreads);
  cfg.dynamicSmemBytes = cp->smem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = kClSize;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  int dbg = 0;
#ifdef AVZ_EXPERIMENT
  if (const char* e = getenv("AVZ_MIX_DBG")) dbg = atoi(e);
#endif
  AVZ_CUDA_OK(cudaLaunchKernelEx(&cfg, k_mix_cluster, src, B, S, cp->st, cp->tw, cp->pairs, prm, peak_eps, vec, dbg, mix, tgt, itf));
  return AVZ_OK;
}

}  // namespace
}  // namespace avz

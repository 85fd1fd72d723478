// Cluster-resident far-field mixer: one thread-block cluster of 8 CTAs mixes one utterance with its spectra never
// leaving the chip.  Included by avz_mixer.cu (same translation unit: MixParams, kMaxSrc live there).
//
// Reference: rt_av_zoom/core/tf_lite_version/world_building.py:46-52 (apply_frac_delay) and :61-93 (mix_and_save) -
// the same arithmetic as the multi-pass path of avz_mixer.cu (pack two real sources per complex transform, per-bin
// ramps, Hermitian re-pack, two inverse transforms, peak normalisation), with a different transform:
//
//   * the two complex planes of an utterance (sources 0+1 and 2+3 on the way in, mic1 + i mic2 and target + i interferer
//     on the way out) live in the shared memory of 2 x 4 CTAs: CTA (p, a) holds the decimated sequence x[4 m + a],
//     m < M = L / 4, of plane p (64 000 samples: 125 KB per CTA);
//   * length-L transform = radix-4 decimation in time: four independent length-M transforms, one per CTA, entirely in
//     its own shared memory (in-place mixed-radix decimation in frequency: radix 16 / 8 / 4 / 2 / 3 / 7 / 5 / 25
//     butterflies in registers, one shared-memory round trip per stage, spectrum left in digit-reversed order), then
//     ONE radix-4 butterfly across the four CTAs per bin: X[k' + M u] = sum_a W4^{a u} W_L^{a k'} Y_a[k'];
//   * that cross-CTA butterfly is not a pass of its own: the thread that owns local bins k' and M - k' reads the 2 x 8
//     values of both planes through distributed shared memory, forms the eight bins k' + M u, (M - k') + M u - a set
//     closed under k -> L - k - does the Hermitian unpack / ramps / re-pack of the four mirror pairs on them, applies
//     the inverse cross-CTA butterfly and writes the 16 values back in place.  Forward cross stage, combine and
//     inverse cross stage cost one DSMEM round trip (DSMEM moves ~17 B/clk per SM - it is what bounds this kernel's
//     exchange steps, so each value crosses it once in and once out);
//   * the inverse local transforms are the exact reverse (conjugate twiddle on the way in, conjugate butterfly, stages
//     backwards: decimation in time) and land in natural order; max|mix| is exchanged through DSMEM, the division by
//     it happens on the way out - no scale pass, no atomics.
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
  int L, M, n;                // n local stages on the M elements of a CTA
  int radix[kClMaxStages];
  int q[kClMaxStages];        // butterfly stride of the stage (block length radix * q)
  int tw_off[kClMaxStages];   // offset of the stage's twiddles [u-1][pos] in the table (stages with q == 1 have none)
  int task_off[kClSize + 1];  // CTA r does tasks [task_off[r], task_off[r + 1]) of the exchange step
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


__device__ constexpr float kW25c[17] = {1.f, 0.96858316112863108f, 0.87630668004386358f, 0.72896862742141155f,
                                        0.53582679497899655f, 0.30901699437494745f, 0.062790519529313527f,
                                        -0.1873813145857246f, -0.42577929156507272f, -0.63742398974868975f,
                                        -0.80901699437494734f, -0.92977648588825135f, -0.99211470131447776f,
                                        -0.99211470131447788f, -0.92977648588825146f, -0.80901699437494778f,
                                        -0.63742398974868952f};
__device__ constexpr float kW25s[17] = {0.f, 0.24868988716485479f, 0.48175367410171532f, 0.68454710592868862f,
                                        0.84432792550201508f, 0.95105651629515353f, 0.99802672842827156f,
                                        0.98228725072868872f, 0.90482705246601947f, 0.77051324277578925f,
                                        0.58778525229247325f, 0.36812455268467814f, 0.12533323356430454f,
                                        -0.12533323356430429f, -0.36812455268467792f, -0.58778525229247269f,
                                        -0.77051324277578936f};

// Two radix-5 stages on 25 consecutive elements in registers (the last two stages of a transform whose length ends in
// 5 x 5): stage q = 5 (inputs v[pos + 5 t]), twiddles W25^{pos u} (constants), stage q = 1 (blocks of five).  The
// inverse runs them backwards with conjugates.  One shared-memory round trip and no twiddle loads for two stages.
template <bool INV>
__device__ __forceinline__ void bfly25(float2 (&v)[25]) {
  if (!INV) {
#pragma unroll
    for (int pos = 0; pos < 5; ++pos) {
      float2 w[5] = {v[pos], v[pos + 5], v[pos + 10], v[pos + 15], v[pos + 20]};
      bfly<5, false>(w);
#pragma unroll
      for (int u = 0; u < 5; ++u)
        v[pos + 5 * u] = (pos * u == 0) ? w[u] : mul_w<false>(w[u], kW25c[pos * u], kW25s[pos * u]);
    }
#pragma unroll
    for (int blk = 0; blk < 5; ++blk) {
      float2 w[5] = {v[5 * blk], v[5 * blk + 1], v[5 * blk + 2], v[5 * blk + 3], v[5 * blk + 4]};
      bfly<5, false>(w);
#pragma unroll
      for (int u = 0; u < 5; ++u) v[5 * blk + u] = w[u];
    }
  } else {
#pragma unroll
    for (int blk = 0; blk < 5; ++blk) {
      float2 w[5] = {v[5 * blk], v[5 * blk + 1], v[5 * blk + 2], v[5 * blk + 3], v[5 * blk + 4]};
      bfly<5, true>(w);
#pragma unroll
      for (int u = 0; u < 5; ++u) v[5 * blk + u] = w[u];
    }
#pragma unroll
    for (int pos = 0; pos < 5; ++pos) {
      float2 w[5];
#pragma unroll
      for (int t = 0; t < 5; ++t)
        w[t] = (pos * t == 0) ? v[pos + 5 * t] : mul_w<true>(v[pos + 5 * t], kW25c[pos * t], kW25s[pos * t]);
      bfly<5, true>(w);
#pragma unroll
      for (int u = 0; u < 5; ++u) v[pos + 5 * u] = w[u];
    }
  }
}

// One stage on a CTA's own M elements: butterflies over x[blk R q + pos + q t], t < R.  QC > 0: the stride is a
// compile-time constant (the stages of the 64 000-sample plan) - every offset becomes an immediate of the load / store
// and the (blk, pos) split a multiply-shift; QC == 0: run-time stride, (blk, pos) advanced incrementally.
// R == 25 stands for the fused pair of radix-5 stages at q = 5 and q = 1 (bfly25).
template <int R, bool INV, int QC>
__device__ __forceinline__ void cl_local_stage(float2* __restrict__ sm, int M, int q_rt, const float2* __restrict__ tw) {
  const int q = QC > 0 ? QC : q_rt;
  const int nbf = M / R;
  int blk = (int)threadIdx.x / q, pos = (int)threadIdx.x - blk * q;
  const int dblk = kClThreads / q, dpos = kClThreads - dblk * q;
  for (int idx = threadIdx.x; idx < nbf; idx += kClThreads) {
    if constexpr (QC > 0) {
      blk = idx / QC;
      pos = idx - blk * QC;
    }
    float2* x = sm + blk * (R * q) + pos;
    float2 v[R];
    if constexpr (R == 25) {
      // The block of 25 is kept in NATURAL order of its bin digit (register d = 5 d1 + d2 holds bin digit d1 + 5 d2:
      // a compile-time transposition).  This digit is the most significant one of the local bin, so the lower half of
      // a mirror pair (k', M - k') always sits in the first 12-13 positions of a block and its partner in the last
      // 12-13 of the mirror block: the exchange step then touches runs of ~100 contiguous bytes of a peer's shared
      // memory instead of isolated pairs (DSMEM moves 32-byte sectors: 2.2x -> 1.25x the payload).
#pragma unroll
      for (int t = 0; t < 25; ++t) v[t] = INV ? x[(t / 5) + 5 * (t % 5)] : x[t];
      bfly25<INV>(v);
#pragma unroll
      for (int u = 0; u < 25; ++u) x[INV ? u : (u / 5) + 5 * (u % 5)] = v[u];
    } else {
#pragma unroll
      for (int t = 0; t < R; ++t) v[t] = x[t * q];
      float2 w[R];
      if (q > 1) {
        if constexpr (R == 16) {
          // W^{pos (4 u2 + u1)} = W^{4 pos u2} W^{pos u1}: six loads and nine products instead of fifteen loads
          float2 w1[4], w4[4];
#pragma unroll
          for (int j = 1; j < 4; ++j) {
            w1[j] = tw[(j - 1) * q + pos];
            w4[j] = tw[(4 * j - 1) * q + pos];
          }
#pragma unroll
          for (int u = 1; u < 16; ++u) {
            const int u1 = u & 3, u2 = u >> 2;
            w[u] = (u2 == 0) ? w1[u1] : (u1 == 0 ? w4[u2] : cmul(w4[u2], w1[u1]));
          }
        } else {
#pragma unroll
          for (int u = 1; u < R; ++u) w[u] = tw[(u - 1) * q + pos];
        }
      }
      if (INV && q > 1) {
#pragma unroll
        for (int t = 1; t < R; ++t) v[t] = cmulc(v[t], w[t]);
      }
      bfly<R, INV>(v);
      if (!INV && q > 1) {
#pragma unroll
        for (int u = 1; u < R; ++u) v[u] = cmul(v[u], w[u]);
      }
#pragma unroll
      for (int u = 0; u < R; ++u) x[u * q] = v[u];
    }
    if (QC == 0) {
      blk += dblk;
      pos += dpos;
      if (pos >= q) {
        pos -= q;
        ++blk;
      }
    }
  }
}

template <bool INV>
__device__ __forceinline__ void cl_local_dispatch(float2* sm, const ClStages& pl, int s, const float2* __restrict__ t) {
  const int R = pl.radix[s], q = pl.q[s], M = pl.M;
  // the 64 000-sample plan (M = 16 000 = 16 x 8 x 5 x 25) with compile-time strides
  if (R == 16 && q == 1000) return cl_local_stage<16, INV, 1000>(sm, M, q, t);
  if (R == 8 && q == 125) return cl_local_stage<8, INV, 125>(sm, M, q, t);
  if (R == 5 && q == 25) return cl_local_stage<5, INV, 25>(sm, M, q, t);
  if (R == 25) return cl_local_stage<25, INV, 1>(sm, M, q, t);
  switch (R) {
    case 16: cl_local_stage<16, INV, 0>(sm, M, q, t); break;
    case 8: cl_local_stage<8, INV, 0>(sm, M, q, t); break;
    case 4: cl_local_stage<4, INV, 0>(sm, M, q, t); break;
    case 2: cl_local_stage<2, INV, 0>(sm, M, q, t); break;
    case 3: cl_local_stage<3, INV, 0>(sm, M, q, t); break;
    case 5: cl_local_stage<5, INV, 0>(sm, M, q, t); break;
    default: cl_local_stage<7, INV, 0>(sm, M, q, t); break;
  }
}

// Hermitian unpack / ramps / re-pack of one mirror pair (k, L - k), k = kk <= L / 2, on both planes, in registers:
// zk[p] = plane p at bin k, zm[p] = plane p at bin L - k (self: the same bin, k = 0 or L / 2).
struct ClCombine {
  const float2* s_ramp;
  int rtl, S, P;
  const int* sym;
  __device__ __forceinline__ void operator()(float2 (&zk)[2], float2 (&zm)[2], int kk, bool self) const {
    float2 m1 = make_float2(0.f, 0.f), m2 = m1, tg = m1;
#pragma unroll
    for (int pp = 0; pp < 2; ++pp) {
      if (pp < P) {
        const float2 av = make_float2(0.5f * (zk[pp].x + zm[pp].x), 0.5f * (zk[pp].y - zm[pp].y));
        const float2 bv = make_float2(0.5f * (zk[pp].y + zm[pp].y), 0.5f * (zm[pp].x - zk[pp].x));
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int s = 2 * pp + h;
          if (s < S) {
            const float2 v = h ? bv : av;
            const float2* t1 = s_ramp + s * rtl;
            const float2 r1 = cmul(t1[kRampLo + (kk >> 8)], t1[kk & (kRampLo - 1)]);
            float2 r2 = make_float2(r1.x, -r1.y);
            if (!sym[s]) {
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
    // x + i y for two real signals x, y: bin k holds X + iY, bin L-k holds conj(X) + i conj(Y)
    zk[0] = make_float2(m1.x - m2.y, m1.y + m2.x);
    zk[1] = make_float2(tg.x - in.y, tg.y + in.x);
    zm[0] = make_float2(m1.x + m2.y, m2.x - m1.y);
    zm[1] = make_float2(tg.x + in.y, in.x - tg.y);
  }
};

// XIO: the 4-way de-interleave between global memory and the decimated sequences runs over DSMEM (below); false: every
// CTA reads / writes its own every-fourth samples directly (4-byte accesses at a 16-byte stride: any alignment).
template <bool XIO>
__global__ void __launch_bounds__(kClThreads, 1)
k_mix_cluster(const float* __restrict__ src, int B, int S, ClStages pl, const float2* __restrict__ tw,
              const int4* __restrict__ tasks, MixParams prm, float peak_eps, int dbg, float* __restrict__ mix,
              float* __restrict__ tgt, float* __restrict__ itf) {
  // dbg (AVZ_EXPERIMENT builds only, 0 otherwise): stop after phase `dbg` and store the planes as they are
  // (1 load, 2 forward local stages, 3 exchange step)
  extern __shared__ __align__(16) unsigned char cl_smem[];
  float2* sm = reinterpret_cast<float2*>(cl_smem);   // Y_a / x_a of this CTA: M complex values
  float2* s_ramp = sm + pl.M;                        // [2 (mic)][4 (source)][kRampLo + n_hi] phase-ramp factors
  float2* s_tw = s_ramp + 2 * 4 * (kRampLo + pl.n_hi);   // twiddles of the late local stages
  __shared__ float s_red[kClThreads / 32];
  __shared__ float s_peak;
  __shared__ int s_sym[kMaxSrc];
  cgx::cluster_group cl = cgx::this_cluster();
  const int rank = (int)cl.block_rank(), p = rank >> 2, a = rank & 3;
  const int n_clusters = gridDim.x / kClSize, cid = blockIdx.x / kClSize;
  const int L = pl.L, M = pl.M;
  const int P = (S + 1) / 2;
  const bool active = 2 * p < S;            // S <= 2: plane 1 has nothing to transform on the way in
  const float inv_n = (float)(1.0 / (double)L);
  const float2* twx = tw + pl.tw_off[kClMaxStages - 1];    // [3][M]: W_L^{a k'(j)}, position order
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
    for (int s = 0; s < pl.n; ++s)
      if (pl.tw_sm[s] >= 0)
        for (int i = threadIdx.x; i < (pl.radix[s] - 1) * pl.q[s]; i += kClThreads)
          s_tw[pl.tw_sm[s] + i] = __ldg(tw + pl.tw_off[s] + i);
    if (threadIdx.x < kMaxSrc) s_sym[threadIdx.x] = prm.sym[threadIdx.x];
    __syncthreads();
  }
  auto stage_tw = [&](int s) -> const float2* { return pl.tw_sm[s] >= 0 ? s_tw + pl.tw_sm[s] : tw + pl.tw_off[s]; };
  const ClCombine combine{s_ramp, kRampLo + pl.n_hi, S, P, s_sym};

  // Global memory <-> decimated sequences.  A CTA's sequence x[4 m + a] is every fourth sample: read directly that is a
  // 4-byte access per 16 bytes, and the four CTAs of a plane each pull every sector of both sources through L2 -> SM
  // (and write quarter sectors on the way out).  XIO: CTA a of a plane instead moves the CONTIGUOUS quarter
  // m in [m_lo, m_hi) of all four sequences - float4 = x[4 m .. 4 m + 3] per source, coalesced, every sector once - and
  // the de-interleave happens over DSMEM: sample j of the vector goes to / comes from CTA j's sm[m] as an 8-byte
  // (re, im) pair, 256 contiguous bytes per warp.  Position m of every peer is touched by exactly one thread of the
  // cluster, so the store of utterance b and the load of utterance b + n_clusters share one loop without a barrier
  // between them.
  const int m_lo = (int)((int64_t)M * a / 4), m_hi = (int)((int64_t)M * (a + 1) / 4);
  float2* pl_sm[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) pl_sm[j] = cl.map_shared_rank(sm, 4 * p + j);
  auto load_x = [&](int b, int m) {
    const float* sa = src + ((int64_t)b * S + 2 * p) * L + 4 * m;
    const float4 va = __ldg(reinterpret_cast<const float4*>(sa));
    const float4 vb = (2 * p + 1 < S) ? __ldg(reinterpret_cast<const float4*>(sa + L)) : make_float4(0.f, 0.f, 0.f, 0.f);
    pl_sm[0][m] = make_float2(va.x, vb.x);
    pl_sm[1][m] = make_float2(va.y, vb.y);
    pl_sm[2][m] = make_float2(va.z, vb.z);
    pl_sm[3][m] = make_float2(va.w, vb.w);
  };
  if constexpr (XIO) {
    if (active && cid < B)
      for (int m = m_lo + threadIdx.x; m < m_hi; m += kClThreads) load_x(cid, m);
    cl_barrier();
  }
  for (int b = cid; b < B; b += n_clusters) {
    // ---- load: two real sources -> re + i im
    if constexpr (!XIO) {
      if (active) {
        const float* sa = src + ((int64_t)b * S + 2 * p) * L + a;
        const bool has_b = 2 * p + 1 < S;
        const float* sb = sa + L;
#pragma unroll 4
        for (int m = threadIdx.x; m < M; m += kClThreads)
          sm[m] = make_float2(__ldg(sa + 4 * m), has_b ? __ldg(sb + 4 * m) : 0.f);
      }
      __syncthreads();
    }
    // ---- forward local stages: Y_a = DFT_M(x_a), digit-reversed order
    if (active && (dbg == 0 || dbg >= 2)) {
      for (int s = 0; s < pl.n; ++s) {
        cl_local_dispatch<false>(sm, pl, s, stage_tw(s));
        __syncthreads();
      }
    }
    // the next utterance's sources: ask L2 for them now, the load at the top of the loop then does not wait for HBM
    if (active && a == 0 && b + n_clusters < B) {
      const char* nx = reinterpret_cast<const char*>(src + ((int64_t)(b + n_clusters) * S + 2 * p) * L);
      const int n_src = (2 * p + 1 < S) ? 2 : 1;
      const int lines = (int)(((size_t)L * sizeof(float) + 127) / 128);
      for (int i = threadIdx.x; i < n_src * lines; i += kClThreads)
        asm volatile("prefetch.global.L2 [%0];" ::"l"(nx + (size_t)i * 128));
    }
    cl_barrier();
    // ---- exchange step: cross-CTA radix-4 butterfly, combine, inverse cross-CTA butterfly - one DSMEM round trip
    const int t_end = (dbg == 0 || dbg >= 3) ? pl.task_off[rank + 1] : 0;
    for (int i = pl.task_off[rank] + threadIdx.x; i < t_end; i += kClThreads) {
      const int4 e = __ldg(tasks + i);
      const int jA = e.x, jB = e.y, kA = e.z;
      const bool two = jA != jB;
      float2 XA[2][4], XB[2][4], wA[4], wB[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
#pragma unroll
        for (int pp = 0; pp < 2; ++pp) {
          XA[pp][u] = make_float2(0.f, 0.f);
          XB[pp][u] = make_float2(0.f, 0.f);
          if (pp < P) {
            const float2* peer = cl.map_shared_rank(sm, 4 * pp + u);
            XA[pp][u] = peer[jA];
            if (two) XB[pp][u] = peer[jB];
          }
        }
        if (u > 0) {
          wA[u] = __ldg(twx + (u - 1) * M + jA);
          wB[u] = __ldg(twx + (u - 1) * M + jB);
        }
      }
#pragma unroll
      for (int pp = 0; pp < 2; ++pp) {
        if (pp < P) {
#pragma unroll
          for (int u = 1; u < 4; ++u) {
            XA[pp][u] = cmul(XA[pp][u], wA[u]);
            XB[pp][u] = cmul(XB[pp][u], wB[u]);
          }
          bfly4<false>(XA[pp][0], XA[pp][1], XA[pp][2], XA[pp][3]);
          bfly4<false>(XB[pp][0], XB[pp][1], XB[pp][2], XB[pp][3]);
        }
      }
      // XA[.][u] = bin kA + M u, XB[.][u] = bin kB + M u, kB = M - kA; mirror of (kA, u) is (kB, 3 - u)
      auto pair = [&](float2 (&X)[2][4], int u, float2 (&Y)[2][4], int v, int kk, bool self) {
        float2 zk[2] = {X[0][u], X[1][u]}, zm[2] = {Y[0][v], Y[1][v]};
        combine(zk, zm, kk, self);
        if (!self) {
          Y[0][v] = zm[0];
          Y[1][v] = zm[1];
        }
        X[0][u] = zk[0];
        X[1][u] = zk[1];
      };
      if (two) {
        const int kB = M - kA;
        pair(XA, 0, XB, 3, kA, false);
        pair(XA, 1, XB, 2, kA + M, false);
        pair(XB, 0, XA, 3, kB, false);
        pair(XB, 1, XA, 2, kB + M, false);
      } else if (kA == 0) {
        pair(XA, 0, XA, 0, 0, true);
        pair(XA, 2, XA, 2, 2 * M, true);
        pair(XA, 1, XA, 3, M, false);
      } else {   // kA == M / 2
        pair(XA, 0, XA, 3, kA, false);
        pair(XA, 1, XA, 2, kA + M, false);
      }
#pragma unroll
      for (int pp = 0; pp < 2; ++pp) {
        bfly4<true>(XA[pp][0], XA[pp][1], XA[pp][2], XA[pp][3]);
        if (two) bfly4<true>(XB[pp][0], XB[pp][1], XB[pp][2], XB[pp][3]);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          float2 za = XA[pp][u], zb = XB[pp][u];
          if (u > 0) {
            za = cmulc(za, wA[u]);
            zb = cmulc(zb, wB[u]);
          }
          float2* peer = cl.map_shared_rank(sm, 4 * pp + u);
          peer[jA] = make_float2(za.x * inv_n, za.y * inv_n);
          if (two) peer[jB] = make_float2(zb.x * inv_n, zb.y * inv_n);
        }
      }
    }
    cl_barrier();
    // ---- inverse local stages, backwards: x_a in natural order, already times 1 / L
    if (dbg == 0) {
      for (int s = pl.n - 1; s >= 0; --s) {
        cl_local_dispatch<true>(sm, pl, s, stage_tw(s));
        __syncthreads();
      }
    }
    // ---- max|mix| over the four CTAs of plane 0
    const bool norm = peak_eps >= 0.f && dbg == 0;
    if (norm && p == 0) {
      float mx = 0.f;
      for (int m = threadIdx.x; m < M; m += kClThreads) {
        const float2 v = sm[m];
        mx = fmaxf(mx, fmaxf(fabsf(v.x), fabsf(v.y)));
      }
      mx = warp_max(mx);
      if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = mx;
      __syncthreads();
      if (threadIdx.x == 0) {
        for (int w = 1; w < kClThreads / 32; ++w) mx = fmaxf(mx, s_red[w]);
        s_peak = mx;
      }
    }
    if (norm || XIO) cl_barrier();   // s_peak of plane 0 (and, XIO, every peer's finished sequence) visible
    // ---- store: / (max|mix| + eps) (world_building.py:86-91)
    float den = 1.f;
    if (norm) {
      float pk = 0.f;
#pragma unroll
      for (int u = 0; u < 4; ++u) pk = fmaxf(pk, *cl.map_shared_rank(&s_peak, u));
      den = pk + peak_eps;
    }
    const float rden = __frcp_rn(den);
    float* o_re = (p == 0 ? mix + (int64_t)b * 2 * L : tgt + (int64_t)b * L);
    float* o_im = (p == 0 ? mix + ((int64_t)b * 2 + 1) * L : itf + (int64_t)b * L);
    if constexpr (XIO) {
      const int bn = b + n_clusters;
      const bool more = active && bn < B;
#pragma unroll 2
      for (int m = m_lo + threadIdx.x; m < m_hi; m += kClThreads) {
        float2 v[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) v[j] = pl_sm[j][m];
        float4 re = make_float4(v[0].x, v[1].x, v[2].x, v[3].x), im = make_float4(v[0].y, v[1].y, v[2].y, v[3].y);
        if (norm) {
          re = make_float4(div_by(re.x, den, rden), div_by(re.y, den, rden), div_by(re.z, den, rden), div_by(re.w, den, rden));
          im = make_float4(div_by(im.x, den, rden), div_by(im.y, den, rden), div_by(im.z, den, rden), div_by(im.w, den, rden));
        }
        *reinterpret_cast<float4*>(o_re + 4 * m) = re;
        *reinterpret_cast<float4*>(o_im + 4 * m) = im;
        if (more) load_x(bn, m);   // the next utterance moves in behind the store
      }
      cl_barrier();   // sequences of the next utterance complete everywhere (and s_peak no longer needed)
    } else {
      o_re += a;
      o_im += a;
#pragma unroll 4
      for (int m = threadIdx.x; m < M; m += kClThreads) {
        const float2 v = sm[m];
        o_re[4 * m] = norm ? div_by(v.x, den, rden) : v.x;
        o_im[4 * m] = norm ? div_by(v.y, den, rden) : v.y;
      }
      __syncthreads();   // everyone has read its sequence before the next utterance's load overwrites it
    }
  }
  cl_barrier();   // nobody leaves while a peer may still be reading its s_peak
}

// ---- host side: plan (radices, twiddle tables, task list of the exchange step) per signal length and device ----
struct ClusterPlan {
  int device;
  int64_t L;
  ClStages st;
  const float2* tw;
  const int4* tasks;
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
  while (m % 16 == 0 && n < kClMaxStages - 1) { st->radix[n++] = 16; m /= 16; }
  for (int r : {8, 4, 2})
    if (m % r == 0 && n < kClMaxStages - 1) { st->radix[n++] = r; m /= r; }
  for (int r : {3, 7, 5})
    while (m % r == 0 && n < kClMaxStages - 1) { st->radix[n++] = r; m /= r; }
  if (m != 1 || n == 0) return false;
  if (n >= 2 && st->radix[n - 1] == 5 && st->radix[n - 2] == 5) {   // ... x 5 x 5: one in-register radix-25 stage
    st->radix[n - 2] = 25;
    --n;
  }
  st->n = n;
  int len = st->M;
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
  const int Li = (int)L, M = st.M;
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
  // the late stages' tables are small and read by every butterfly of every utterance: they get a copy in shared memory
  st.tw_sm_n = 0;
  for (int s = 0; s < kClMaxStages; ++s) st.tw_sm[s] = -1;
  for (int s = st.n - 1; s >= 0; --s) {
    const int n = (st.radix[s] - 1) * st.q[s];
    if (st.q[s] == 1) continue;
    if (st.tw_sm_n + n > kTwSmemMax) break;
    st.tw_sm[s] = st.tw_sm_n;
    st.tw_sm_n += n;
  }
  // local bin of every position (digits reversed; a radix-25 stage is the digit pair of its two radix-5 stages)
  std::vector<int> bin_of((size_t)M), pos_of((size_t)M);
  for (int pos = 0; pos < M; ++pos) {
    int k = 0, mult = 1;
    for (int s = 0; s < st.n; ++s) {
      const int d = (pos / st.q[s]) % st.radix[s];
      k += d * mult;   // (a radix-25 block is stored in natural order of its digit - cl_local_stage)
      mult *= st.radix[s];
    }
    bin_of[(size_t)pos] = k;
    pos_of[(size_t)k] = pos;
  }
  // cross-CTA twiddles W_L^{a k'} in position order
  st.tw_off[kClMaxStages - 1] = (int)tw.size();
  for (int a = 1; a < 4; ++a)
    for (int pos = 0; pos < M; ++pos) {
      const double ang = two_pi * (double)(((int64_t)a * bin_of[(size_t)pos]) % Li) / (double)Li;
      tw.push_back(make_float2((float)cos(ang), (float)-sin(ang)));
    }
  // tasks of the exchange step: local bins (k', M - k'), k' <= M / 2, as positions, sorted by the first position; each
  // CTA of the cluster takes an eighth
  std::vector<int4> tasks;
  for (int pos = 0; pos < M; ++pos) {
    const int k = bin_of[(size_t)pos];
    const int kb = (M - k) % M;
    if (k > kb) continue;
    tasks.push_back(make_int4(pos, pos_of[(size_t)kb], k, 0));
  }
  for (int r = 0; r <= kClSize; ++r) st.task_off[r] = (int)((int64_t)tasks.size() * r / kClSize);
  void *dtw = nullptr, *dts = nullptr;
  AVZ_CUDA_OK(cudaMalloc(&dtw, tw.size() * sizeof(float2)));
  AVZ_CUDA_OK(cudaMalloc(&dts, tasks.size() * sizeof(int4)));
  AVZ_CUDA_OK(cudaMemcpy(dtw, tw.data(), tw.size() * sizeof(float2), cudaMemcpyHostToDevice));
  AVZ_CUDA_OK(cudaMemcpy(dts, tasks.data(), tasks.size() * sizeof(int4), cudaMemcpyHostToDevice));
  ClusterPlan* p = new ClusterPlan{dev, L, st, (const float2*)dtw, (const int4*)dts, 0,
                                   (size_t)M * sizeof(float2) + cl_ramp_bytes(st.n_hi) + (size_t)st.tw_sm_n * sizeof(float2)};
  // how many such clusters the device runs at once (0: it cannot - the caller falls back to the multi-pass path)
  if (cudaFuncSetAttribute(k_mix_cluster<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p->smem) == cudaSuccess &&
      cudaFuncSetAttribute(k_mix_cluster<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p->smem) == cudaSuccess) {
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
    int nc2 = 0;
    if (cudaOccupancyMaxActiveClusters(&nc, k_mix_cluster<true>, &cfg) == cudaSuccess &&
        cudaOccupancyMaxActiveClusters(&nc2, k_mix_cluster<false>, &cfg) == cudaSuccess)
      p->max_clusters = nc < nc2 ? nc : nc2;
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
    AVZ_CUDA_OK(cudaFuncSetAttribute(k_mix_cluster<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cp->smem));
    AVZ_CUDA_OK(cudaFuncSetAttribute(k_mix_cluster<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cp->smem));
  }
  const int n_clusters = B < cp->max_clusters ? B : cp->max_clusters;
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
  // float4 traffic needs 16-byte aligned signals (L is a multiple of 4 here, so every utterance then is)
  bool xio = (((uintptr_t)src | (uintptr_t)mix | (uintptr_t)tgt | (uintptr_t)itf) & 15) == 0;
#ifdef AVZ_EXPERIMENT
  if (const char* e = getenv("AVZ_MIX_XIO")) xio = xio && atoi(e) != 0;
#endif
  if (xio) {
    AVZ_CUDA_OK(cudaLaunchKernelEx(&cfg, k_mix_cluster<true>, src, B, S, cp->st, cp->tw, cp->tasks, prm, peak_eps, dbg, mix, tgt, itf));
  } else {
    AVZ_CUDA_OK(cudaLaunchKernelEx(&cfg, k_mix_cluster<false>, src, B, S, cp->st, cp->tw, cp->tasks, prm, peak_eps, dbg, mix, tgt, itf));
  }
  return AVZ_OK;
}

}  // namespace
}  // namespace avz

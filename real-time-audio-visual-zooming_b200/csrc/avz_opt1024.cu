// n_fft = 1024 / hop 512 fast path - the STFT shape of the reference's learned pipelines
// (rt_av_zoom/core/full_audio_generating_pipeline/inference.py:88-117, tf_lite_version/inference.py:85-179,307-349,
// Final_pipeline/src/inference.py:198-222 with their config.json) - on the register-resident 512-point warp FFT.
//
// A real 1024-sample frame x is transformed as ONE 512-point complex FFT of z[n] = x[2n] + i x[2n+1]:
//   with A = Z[k] + conj Z[512-k],  T = W1024^k * (-i) (Z[k] - conj Z[512-k]):
//   X[k] = (A + T) / 2,   X[512-k] = conj(A - T) / 2          (k = 0..255; X[256] = conj Z[256])
// so lane (k1, h) of avz_fft512.cuh, which holds lo bins k = k1 + 16 j + 128 h and (after one shuffle per bin) their
// mirrors Z[512-k], produces bins k and 512-k of the one-sided spectrum.  The inverse runs the same algebra backwards
// (E = S[k] + conj S[512-k], O = (S[k] - conj S[512-k]) conj W1024^k, Z = (E + i O) / 2) and one inverse 512-point
// transform returns the frame as (even, odd) sample pairs, which is also the layout of coalesced 8-byte loads/stores.
//
//   k1024_features<PHYS>  STFT(mic0), STFT(mic1) -> ln(|Y0| + 1e-7), angle(Y0) - angle(Y1); NCHW rows leave as float4
//                         runs along T (PHYS: the 4-channel NHWC layout, stored from the bin loop)
//   k1024_cov<KEEP>       STFT of both mics -> (1 - mask)-weighted 2x2 covariance partial sums, accumulators in
//                         registers (KEEP: both spectra are also streamed out for pass B)
//   k1024_apply<KEPT>     STFT of both mics (or the kept spectra) -> w^H y -> post-filter gain -> inverse -> window ->
//                         overlap-add of the two half-frames in registers -> / sum w^2 -> coalesced stores (+ peak)
//
// The two per-channel spectra of a frame meet in a per-warp shared-memory buffer in natural bin order (513 + 513
// complex): the covariance / beamforming arithmetic then runs on bins lane + 32 i, where the mask tile, the weights and
// the partial sums are contiguous.  One warp owns a run of consecutive frames; the frame path has only __syncwarp()
// (plus one block barrier per CTA when the asynchronously staged mask tile is first used).  Measurements and the ncu
// findings that shaped this file: profiles/r1_ncu_1024.md.
#include <cstdlib>

#include "avz_common.cuh"
#include "avz_fft512.cuh"

namespace avz {
namespace o1024 {

using f512::Lane;

constexpr int kN = 1024;
constexpr int kHop = 512;
constexpr int kF = 513;
constexpr int kFP = 544;      // padded bins of the partial sums (Geo<1024>::FP)
constexpr int kBPL = 17;      // bins lane + 32 i, i < 17, k <= 512
constexpr int kWarps = 4;
constexpr int kYP = 520;      // complex elements per channel plane of the per-warp spectrum buffer
constexpr int kFeatFrames = 8;
constexpr int kFeatPitch = kFeatFrames + 1;

constexpr int kTileFrames = 16;   // frames per CTA of the mask-reading kernels: their mask tile lives in shared memory
constexpr int kTilePitch = kTileFrames + 1;

enum { GAIN_NONE = 0, GAIN_BITS = 1, GAIN_FLOOR = 2, GAIN_MASK = 3 };   // as avz_generic.cu

// Where the two channels of utterance b live and how many of its L samples exist.  A plain batch is [B][2][L]
// (n_win = 1, rec_stride = 2 L, ch_stride = L, rec_len = L).  The chunk drivers of the reference
// (full_audio.../inference.py:137-147, Final_pipeline/src/inference.py:186-193) cut a recording into windows of L
// samples at stride win_stride, zero-padded past its end: here utterance b = r * n_win + i is read IN PLACE from
// recording r of a planar [R][2][rec_len] buffer - no gathered copy of the windows exists.
struct WaveView {
  int64_t rec_stride, ch_stride;
  int n_win, win_stride, rec_len;
  __device__ __forceinline__ const float* chan0(const float* base, int b, int L, int& valid) const {
    const int r = b / n_win, i = b - r * n_win;
    const int start = i * win_stride;
    valid = max(0, min(L, rec_len - start));
    return base + (int64_t)r * rec_stride + start;
  }
};

// The caller's mask is (B, F, T): one frame's 513 values are 513 different sectors.  Each CTA therefore copies the
// (513 x <= 16 frames) tile it needs into shared memory once, asynchronously (cp.async, 16 consecutive threads per
// 64-byte row), and every frame then reads its weights with consecutive lanes on consecutive bins (pitch 17: no bank
// conflicts).  The copy overlaps the first frame's transforms; tile_ready() is the wait + barrier before first use.
__device__ __forceinline__ void stage_mask_tile(float* __restrict__ s_tile, const float* __restrict__ mk, int T, int t0,
                                                int nt) {
  for (int idx = threadIdx.x; idx < kF * kTileFrames; idx += kWarps * 32) {
    const int k = idx / kTileFrames, tl = idx - k * kTileFrames;
    if (tl < nt)
      asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(s_tile + k * kTilePitch + tl)),
                   "l"(mk + (int64_t)k * T + t0 + tl)
                   : "memory");
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
}
__device__ __forceinline__ void tile_ready() {
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();
}

#ifndef AVZ_MINB_1024_COV
#define AVZ_MINB_1024_COV 2
#endif
#ifndef AVZ_1024_FEAT_UNROLL
#define AVZ_1024_FEAT_UNROLL 8
#endif
#ifndef AVZ_MINB_1024_APPLY
#define AVZ_MINB_1024_APPLY 2
#endif

struct Ctx {
  Lane ln;
  float2 wk[8];   // W1024^k for this lane's lo bins k = k1 + 16 j + 128 h
  __device__ __forceinline__ void init(const float2* __restrict__ tw512, const float2* __restrict__ tw1024) {
    ln.init(tw512);
#pragma unroll
    for (int j = 0; j < 8; ++j) wk[j] = tw1024[ln.k1 + 16 * j + 128 * ln.h];
  }
};

// Shared-memory constants of a CTA: analysis window (w[2n], w[2n+1]) * lane sign / 1024 and synthesis window
// (w[2n], w[2n+1]) * lane sign / 2, both indexed by n = 32 r + lane (the transform's time layout).
__device__ __forceinline__ void fill_windows(float2* s_wa, float2* s_ws, const float* __restrict__ win) {
  for (int n = threadIdx.x; n < 512; n += blockDim.x) {
    const float sg = ((n & 3) == 3) ? -1.f : 1.f;
    const float a = win[2 * n], b = win[2 * n + 1];
    s_wa[n] = make_float2(a * sg * (1.f / 1024.f), b * sg * (1.f / 1024.f));
    if (s_ws) s_ws[n] = make_float2(a * sg * 0.5f, b * sg * 0.5f);
  }
}

// Raw samples of frame t of one channel: raw[r] = (x[s + 64 r + 2 lane], x[... + 1]), s = 512 t - 512; zero outside
// [0, L) (scipy boundary='zeros', padded=True).
__device__ __forceinline__ void load_frame(float2 (&raw)[16], const float* __restrict__ x, int L, int t, int lane) {
  const int s = t * kHop - kN / 2;
  if (s >= 0 && s + kN <= L && (reinterpret_cast<uintptr_t>(x) & 7) == 0) {   // warp-uniform
    const float2* p = reinterpret_cast<const float2*>(x + s) + lane;
#pragma unroll
    for (int r = 0; r < 16; ++r) raw[r] = __ldg(p + 32 * r);
  } else {
#pragma unroll
    for (int r = 0; r < 16; ++r) {
      const int i = s + 64 * r + 2 * lane;
      raw[r].x = ((unsigned)i < (unsigned)L) ? __ldg(x + i) : 0.f;
      raw[r].y = ((unsigned)(i + 1) < (unsigned)L) ? __ldg(x + i + 1) : 0.f;
    }
  }
}

// One channel: window, 512-point transform, even/odd split.  lo[j] = X[k], up[j] = X[512 - k] for this lane's lo bins
// (scaled by 1 / sum(w) as scipy's stft); mid = X[256] (meaningful on lane 0 only).
__device__ __forceinline__ void analyse(const float2 (&raw)[16], const float2* __restrict__ s_wa, float2* __restrict__ sm,
                                        const Ctx& cx, float2 (&lo)[8], float2 (&up)[8], float2& mid) {
  float2 v[16];
#pragma unroll
  for (int r = 0; r < 16; ++r) {
    const float2 w = s_wa[32 * r + cx.ln.lane];
    v[r] = make_float2(raw[r].x * w.x, raw[r].y * w.y);
  }
  f512::forward(v, sm, cx.ln);
  float2 mir[8];
  f512::mirror_of_low(v, mir, cx.ln);
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float2 z = v[j], m = mir[j];
    const float2 A = make_float2(z.x + m.x, z.y - m.y);          // Z[k] + conj Z[512-k]
    const float2 Bv = make_float2(z.y + m.y, m.x - z.x);         // -i (Z[k] - conj Z[512-k])
    const float2 T = cmul(Bv, cx.wk[j]);
    lo[j] = make_float2(A.x + T.x, A.y + T.y);
    up[j] = make_float2(A.x - T.x, T.y - A.y);                    // conj(A - T)
  }
  mid = make_float2(2.f * v[8].x, -2.f * v[8].y);                // lane 0: hi[0] = Z[256]; X[256] = conj Z[256]
}

// ... and into the per-warp buffer in natural bin order.
__device__ __forceinline__ void spectrum_to_smem(float2* __restrict__ Yc, const float2 (&lo)[8], const float2 (&up)[8],
                                                 float2 mid, const Lane& ln) {
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int k = ln.k1 + 16 * j + 128 * ln.h;
    Yc[k] = lo[j];
    Yc[512 - k] = up[j];
  }
  if (ln.lane == 0) Yc[256] = mid;
}

// Both channels of a frame into Y0 / Y1.  The channel loop is deliberately NOT unrolled: one copy of the transform in
// the instruction stream instead of two (the first version of these kernels stalled on instruction fetch - ncu
// "no_instruction" 0.5 to 4.9 stalls per issue - because three to four inlined 512-point transforms do not fit the
// instruction cache).  ra holds channel 0 on entry and is free afterwards; rb holds channel 1.
__device__ __forceinline__ void analyse_pair(float2 (&ra)[16], const float2 (&rb)[16], const float2* __restrict__ s_wa,
                                             float2* __restrict__ sm, const Ctx& cx, float2* __restrict__ Y0,
                                             float2* __restrict__ Y1) {
#pragma unroll 1
  for (int ch = 0; ch < 2; ++ch) {
    float2 lo[8], up[8], mid;
    analyse(ra, s_wa, sm, cx, lo, up, mid);
    spectrum_to_smem(ch ? Y1 : Y0, lo, up, mid, cx.ln);
#pragma unroll
    for (int r = 0; r < 16; ++r) ra[r] = rb[r];
  }
}

// ------------------------------------------------------------------------------------------
// features
// ------------------------------------------------------------------------------------------
// Persistent CTAs loop over (utterance, 8-frame tile) units: windows and lane constants are set up once per CTA.
// PHYS: the 4-channel NHWC layout of Final_pipeline/src/inference.py:117-128 (stored straight from the bin loop; its
// sincosf lives only in that instantiation).
template <bool PHYS>
__device__ __forceinline__ void feature_bin(const float2* __restrict__ Y0, const float2* __restrict__ Y1, int k,
                                            float* __restrict__ X, float* __restrict__ s_tile, int b, int t, int tl, int T) {
  float lm, ipd;
  feature_values(Y0[k], Y1[k], lm, ipd);
  if (PHYS) {
    store_features(X, AVZ_FEAT_PHYSICS_NHWC, b, k, t, kF, T, lm, ipd);
  } else {
    s_tile[k * kFeatPitch + tl] = lm;
    s_tile[(kF + k) * kFeatPitch + tl] = ipd;
  }
}

template <bool PHYS>
__global__ void __launch_bounds__(kWarps * 32, 2)
k1024_features(const float* __restrict__ mix, WaveView vw, int L, int T, int B, int mode, float* __restrict__ X,
               Tables tb512, Tables tb) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float2* s_wa = reinterpret_cast<float2*>(smem_raw);                        // [512]
  float2* s_fft = s_wa + 512;                                                // [kWarps][kSmemComplex]
  float2* s_y = s_fft + kWarps * f512::kSmemComplex;                         // [kWarps][2][kYP]
  float* s_tile = reinterpret_cast<float*>(s_y + kWarps * 2 * kYP);          // [2][kF][kFeatPitch]
  fill_windows(s_wa, nullptr, tb.win);
  __syncthreads();
  Ctx cx;
  cx.init(tb512.tw, tb.tw);
  const int lane = cx.ln.lane, warp = warp_id_uniform();
  float2* sm = s_fft + (size_t)warp * f512::kSmemComplex;
  float2* Y0 = s_y + (size_t)warp * 2 * kYP;
  float2* Y1 = Y0 + kYP;
  const bool wrapped = (mode == AVZ_FEAT_LOGMAG_IPD_WRAPPED);
  const int tiles_per_utt = (T + kFeatFrames - 1) / kFeatFrames;
  const int n_tiles = B * tiles_per_utt;
#pragma unroll 1
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int b = tile / tiles_per_utt;
    const int t0 = (tile - b * tiles_per_utt) * kFeatFrames;
    const int nt = min(kFeatFrames, T - t0);
    int Lv;
    const float* m0 = vw.chan0(mix, b, L, Lv);
    const float* m1 = m0 + vw.ch_stride;
#pragma unroll 1
    for (int tl = warp; tl < nt; tl += kWarps) {
      const int t = t0 + tl;
      float2 r0[16], r1[16];
      load_frame(r0, m0, Lv, t, lane);
      load_frame(r1, m1, Lv, t, lane);
      analyse_pair(r0, r1, s_wa, sm, cx, Y0, Y1);
      __syncwarp();
#pragma unroll 1
      for (int i0 = 0; i0 < 16; i0 += AVZ_1024_FEAT_UNROLL) {   // independent bins in flight: the per-bin chain is long
#pragma unroll
        for (int u = 0; u < AVZ_1024_FEAT_UNROLL; ++u)
          feature_bin<PHYS>(Y0, Y1, lane + 32 * (i0 + u), X, s_tile, b, t, tl, T);
      }
      if (lane == 0) feature_bin<PHYS>(Y0, Y1, 512, X, s_tile, b, t, tl, T);
      __syncwarp();
    }
    if (PHYS) continue;
    __syncthreads();
    // rows (feature, bin) leave as runs along T.  Full tiles with 16-byte-aligned rows: two threads per row, one
    // float4 each (a warp writes 16 whole sectors per instruction; the scalar loop below cost 800 instructions per
    // frame - a fifth of the kernel); ragged tiles / odd T: kFeatFrames consecutive threads per row.
    const bool vec = (nt == kFeatFrames) && ((T & 3) == 0) && ((reinterpret_cast<uintptr_t>(X) & 15) == 0);
    static_assert(kFeatFrames == 8, "the vector store path assumes two float4 per row");
    if (vec) {
      for (int idx = threadIdx.x; idx < 2 * kF * 2; idx += kWarps * 32) {
        const int row = idx >> 1, half = idx & 1;
        const float* src = s_tile + row * kFeatPitch + 4 * half;
        float4 v = make_float4(src[0], src[1], src[2], src[3]);
        if (wrapped && row >= kF) {
          const float two_pi = 6.28318530717958647692f;
          v.x -= two_pi * rintf(v.x / two_pi);
          v.y -= two_pi * rintf(v.y / two_pi);
          v.z -= two_pi * rintf(v.z / two_pi);
          v.w -= two_pi * rintf(v.w / two_pi);
        }
        *reinterpret_cast<float4*>(X + ((int64_t)b * 2 * kF + row) * T + t0 + 4 * half) = v;
      }
    }
    for (int idx = vec ? 2 * kF * kFeatFrames : threadIdx.x; idx < 2 * kF * kFeatFrames; idx += kWarps * 32) {
      const int row = idx / kFeatFrames, tl = idx - row * kFeatFrames;
      if (tl < nt) {
        float v = s_tile[(size_t)row * kFeatPitch + tl];
        if (wrapped && row >= kF) {
          const float two_pi = 6.28318530717958647692f;
          v = v - two_pi * rintf(v / two_pi);
        }
        X[((int64_t)b * 2 * kF + row) * T + t0 + tl] = v;
      }
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------
// mask-weighted covariance partial sums: part[B][chunks][5][kFP] = (R00, R11, Re R01, Im R01, sum m), un-normalised
// ------------------------------------------------------------------------------------------
// KEEP: both one-sided spectra of every frame are also written to `spec` ([B][T][2][kYP] complex, natural bin order,
// 8320 B per frame = 16 B per sample) so that pass B starts from them instead of transforming the waveform again.
template <bool KEEP>
__global__ void __launch_bounds__(kWarps * 32, AVZ_MINB_1024_COV)
k1024_cov(const float* __restrict__ mix, WaveView vw, const float* __restrict__ mask, int L_full, int T, int frames_per_cta,
          float sqrt_eps, float* __restrict__ part, float2* __restrict__ spec, Tables tb512, Tables tb) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float2* s_wa = reinterpret_cast<float2*>(smem_raw);                        // [512]
  float2* s_fft = s_wa + 512;                                                // [kWarps][kSmemComplex]
  float2* s_y = s_fft + kWarps * f512::kSmemComplex;                         // [kWarps][2][kYP]
  float* s_tile = reinterpret_cast<float*>(s_y + kWarps * 2 * kYP);          // [kF][kTilePitch]
  const int b = blockIdx.y, chunk = blockIdx.x, chunks = gridDim.x;
  const int c0 = chunk * frames_per_cta, c1 = min(T, c0 + frames_per_cta);   // frames_per_cta <= kTileFrames
  stage_mask_tile(s_tile, mask + (int64_t)b * kF * T, T, c0, c1 - c0);
  fill_windows(s_wa, nullptr, tb.win);
  __syncthreads();
  Ctx cx;
  cx.init(tb512.tw, tb.tw);
  const int lane = cx.ln.lane, warp = warp_id_uniform();
  float2* sm = s_fft + (size_t)warp * f512::kSmemComplex;
  float2* Y0 = s_y + (size_t)warp * 2 * kYP;
  float2* Y1 = Y0 + kYP;
  int L;
  const float* m0 = vw.chan0(mix, b, L_full, L);
  const float* m1 = m0 + vw.ch_stride;
  bool tile_pending = true;

  const int per = (c1 - c0 + kWarps - 1) / kWarps;
  const int ta = c0 + warp * per, tb_ = min(c1, ta + per);

  float acc[5][kBPL];
#pragma unroll
  for (int q = 0; q < 5; ++q)
#pragma unroll
    for (int i = 0; i < kBPL; ++i) acc[q][i] = 0.f;

  float2 r0[16];
  if (ta < tb_) load_frame(r0, m0, L, ta, lane);
#pragma unroll 1
  for (int t = ta; t < tb_; ++t) {
    float2 r1[16];
    load_frame(r1, m1, L, t, lane);
    analyse_pair(r0, r1, s_wa, sm, cx, Y0, Y1);
    if (t + 1 < tb_) load_frame(r0, m0, L, t + 1, lane);   // in flight during the accumulation below
    if (tile_pending) {
      tile_ready();
      tile_pending = false;
    }
    float wgt[kBPL];
#pragma unroll
    for (int i = 0; i < kBPL; ++i) {
      const int k = lane + 32 * i;
      wgt[i] = (k <= 512) ? 1.f - s_tile[k * kTilePitch + (t - c0)] : 0.f;
    }
    __syncwarp();
#pragma unroll
    for (int i = 0; i < kBPL; ++i) {
      const int k = lane + 32 * i;
      if (k <= 512) {
        const float2 y0 = Y0[k], y1 = Y1[k];
        if (KEEP) {
          float2* sp = spec + ((int64_t)b * T + t) * 2 * kYP;
          __stcs(sp + k, y0);
          __stcs(sp + kYP + k, y1);
        }
        const float m = wgt[i];
        const float ms = m + sqrt_eps;
        const float2 c01 = cmulc(y0, y1);
        acc[0][i] = fmaf(ms, cabs2(y0), acc[0][i]);
        acc[1][i] = fmaf(ms, cabs2(y1), acc[1][i]);
        acc[2][i] = fmaf(ms, c01.x, acc[2][i]);
        acc[3][i] = fmaf(ms, c01.y, acc[3][i]);
        acc[4][i] += m;
      }
    }
    __syncwarp();
  }
  if (tile_pending) tile_ready();   // warps without frames still take part in the barrier
  // fixed-order reduction over the CTA's warps (bit-stable reruns); the staging area reuses the frame buffers
  __syncthreads();
  float* s_acc = reinterpret_cast<float*>(s_fft);   // [kWarps][5][kFP]
#pragma unroll
  for (int q = 0; q < 5; ++q)
#pragma unroll
    for (int i = 0; i < kBPL; ++i) s_acc[((size_t)warp * 5 + q) * kFP + lane + 32 * i] = acc[q][i];
  __syncthreads();
  float* dst = part + ((int64_t)b * chunks + chunk) * 5 * kFP;
  for (int i = threadIdx.x; i < 5 * kFP; i += kWarps * 32) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) s += s_acc[(size_t)w * 5 * kFP + i];
    dst[i] = s;
  }
}

// ------------------------------------------------------------------------------------------
// beamform + post-filter + inverse + overlap-add
// ------------------------------------------------------------------------------------------
// KEPT: the spectra come from `spec` (written by k1024_cov<true>) instead of two forward transforms; the arithmetic
// after that point is the same code, so both variants give bit-identical waveforms.
template <bool KEPT>
__global__ void __launch_bounds__(kWarps * 32, AVZ_MINB_1024_APPLY)
k1024_apply(const float* __restrict__ mix, WaveView vw, const float2* __restrict__ spec, const float2* __restrict__ w,
            const float* __restrict__ mask, int gain_mode, float post_floor, int L_full, int T, int blocks_per_cta,
            float* __restrict__ out, float* __restrict__ peak, Tables tb512, Tables tb) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float2* s_wa = reinterpret_cast<float2*>(smem_raw);                        // [512]
  float2* s_ws = s_wa + 512;                                                 // [512]
  float2* s_cw = s_ws + 512;                                                 // [2][kYP]: conj(w0), conj(w1)
  float2* s_fft = s_cw + 2 * kYP;                                            // [kWarps][kSmemComplex]
  float2* s_y = s_fft + kWarps * f512::kSmemComplex;                         // [kWarps][2][kYP]
  float* s_tile = reinterpret_cast<float*>(s_y + kWarps * 2 * kYP);          // [kF][kTilePitch]
  __shared__ float s_peak[kWarps];
  const int b = blockIdx.y;
  const bool use_mask = (gain_mode == GAIN_FLOOR || gain_mode == GAIN_MASK);
  {
    const int g0 = 1 + blockIdx.x * blocks_per_cta, g1 = min(T, g0 + blocks_per_cta);   // blocks_per_cta < kTileFrames
    if (use_mask) stage_mask_tile(s_tile, mask + (int64_t)b * kF * T, T, g0 - 1, g1 - g0 + 1);
  }
  bool tile_pending = use_mask;
  fill_windows(s_wa, s_ws, tb.win);
  for (int k = threadIdx.x; k < kF; k += kWarps * 32) {
    const float2 w0 = w[((int64_t)b * kF + k) * 2 + 0];
    const float2 w1 = w[((int64_t)b * kF + k) * 2 + 1];
    s_cw[k] = make_float2(w0.x, -w0.y);
    s_cw[kYP + k] = make_float2(w1.x, -w1.y);
  }
  __syncthreads();
  Ctx cx;
  cx.init(tb512.tw, tb.tw);
  const int lane = cx.ln.lane, warp = warp_id_uniform();
  float2* sm = s_fft + (size_t)warp * f512::kSmemComplex;
  float2* Y0 = s_y + (size_t)warp * 2 * kYP;
  float2* Y1 = Y0 + kYP;
  int L = L_full;
  const float* m0 = KEPT ? nullptr : vw.chan0(mix, b, L_full, L);
  const float* m1 = KEPT ? nullptr : m0 + vw.ch_stride;
  float* ob = out + (int64_t)b * (int64_t)(T - 1) * kHop;

  // Output block g (1 <= g <= T-1) = second half of frame g-1 + first half of frame g, stored at (g-1) * 512.
  // The CTA emits blocks [G0, G1), i.e. it transforms frames G0-1 .. G1-1, split into one run of consecutive frames
  // per warp.  Inside a run the open half-frame stays in registers.  The first block of a run needs the last
  // half-frame of the previous warp's run: the warp parks its own half in the output buffer, the previous warp leaves
  // its tail in shared memory (in its spectrum buffer, dead by then), and the sum is completed after one barrier.
  // Only the CTA's very first frame (G0-1) is a warm-up whose first half is discarded.
  const int G0 = 1 + blockIdx.x * blocks_per_cta, G1 = min(T, G0 + blocks_per_cta);
  const int F0 = G0 - 1, nf = G1 - G0 + 1;
  const int per = (nf + kWarps - 1) / kWarps;
  const int fa = F0 + warp * per, fb = min(F0 + nf, fa + per);

  float2 tail[8];
  float my_peak = 0.f;
  float2 r0[16];                 // !KEPT: channel 0 of the frame to come
  float2 ya[kBPL], yb[kBPL];     // KEPT: both spectra of the frame to come, bins lane + 32 i
  auto load_kept = [&](int t) {
    const float2* sp = spec + ((int64_t)b * T + t) * 2 * kYP + lane;
#pragma unroll
    for (int i = 0; i < kBPL; ++i) {
      const bool ok = (i < kBPL - 1) || lane == 0;
      ya[i] = ok ? __ldcs(sp + 32 * i) : make_float2(0.f, 0.f);
      yb[i] = ok ? __ldcs(sp + kYP + 32 * i) : make_float2(0.f, 0.f);
    }
  };
  if (fa < fb) {
    if (KEPT) load_kept(fa);
    else load_frame(r0, m0, L, fa, lane);
  }
#pragma unroll 1
  for (int t = fa; t < fb; ++t) {
    if (!KEPT) {
      float2 r1[16];
      load_frame(r1, m1, L, t, lane);
      analyse_pair(r0, r1, s_wa, sm, cx, Y0, Y1);
      if (t + 1 < fb) load_frame(r0, m0, L, t + 1, lane);
    }
    if (tile_pending) {
      tile_ready();
      tile_pending = false;
    }
    __syncwarp();
    // S[k] = conj(w0) Y0 + conj(w1) Y1, times the post-filter gain; the c2r transform ignores Im(DC), Im(Nyquist)
#pragma unroll
    for (int i = 0; i < kBPL; ++i) {
      const int k = lane + 32 * i;
      if (k <= 512) {
        float g = 1.f;
        if (use_mask) {
          g = s_tile[k * kTilePitch + (t - F0)];
          if (gain_mode == GAIN_FLOOR) g = fmaxf(g, post_floor);
        }
        const float2 y0 = KEPT ? ya[i] : Y0[k], y1 = KEPT ? yb[i] : Y1[k];
        const float2 s = cadd(cmul(s_cw[k], y0), cmul(s_cw[kYP + k], y1));
        Y0[k] = make_float2(s.x * g, (k == 0 || k == 512) ? 0.f : s.y * g);
      }
    }
    if (KEPT && t + 1 < fb) load_kept(t + 1);   // in flight during the inverse transform below
    if (KEPT && t + 2 < fb) {                    // and the frame after that on its way from DRAM to L2 (no registers)
      const char* nx = reinterpret_cast<const char*>(spec + ((int64_t)b * T + t + 2) * 2 * kYP);
#pragma unroll
      for (int q = 0; q < 3; ++q) {
        const int off = (lane + 32 * q) * 128;
        if (off < 2 * kYP * (int)sizeof(float2)) asm volatile("prefetch.global.L2 [%0];" ::"l"(nx + off));
      }
    }
    __syncwarp();
    float2 v[16];
    {
      float2 Sa[8], Sb[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int k = cx.ln.k1 + 16 * j + 128 * cx.ln.h;
        const float2 a = Y0[k], c = Y0[512 - k];
        Sa[j] = make_float2(a.x + c.x, a.y - c.y);                               // S[k] + conj S[512-k]
        Sb[j] = cmulc(make_float2(a.x - c.x, a.y + c.y), cx.wk[j]);              // (S[k] - conj S[512-k]) conj W^k
      }
      const float2 s256 = Y0[256];
      f512::hermitian_pack(Sa, Sb, make_float2(2.f * s256.x, -2.f * s256.y), v, cx.ln);
    }
    __syncwarp();
    f512::inverse(v, sm, cx.ln);
    // v[r] * s_ws = irfft(S) * sum(w) * w  at samples 64 r + 2 lane, + 1
    if (t > F0) {
      float2* dst = reinterpret_cast<float2*>(ob + (int64_t)(t - 1) * kHop) + lane;
      if (t == fa) {   // first frame of a later warp's run: park the half-frame, completed after the barrier
#pragma unroll
        for (int r = 0; r < 8; ++r) {
          const float2 wa = s_ws[32 * r + lane];
          dst[32 * r] = make_float2(v[r].x * wa.x, v[r].y * wa.y);
        }
      } else {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
          const float2 wa = s_ws[32 * r + lane], wb = s_ws[32 * (r + 8) + lane];   // |.| = w / 2 (sign folded in)
          const float2 cur = make_float2(v[r].x * wa.x, v[r].y * wa.y);
          const float nx = 4.f * fmaf(wa.x, wa.x, wb.x * wb.x), ny = 4.f * fmaf(wa.y, wa.y, wb.y * wb.y);
          float2 o;
          o.x = (tail[r].x + cur.x) / (nx > 1e-10f ? nx : 1.f);
          o.y = (tail[r].y + cur.y) / (ny > 1e-10f ? ny : 1.f);
          dst[32 * r] = o;
          my_peak = fmaxf(my_peak, fmaxf(fabsf(o.x), fabsf(o.y)));
        }
      }
    }
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      const float2 wb = s_ws[32 * (r + 8) + lane];
      tail[r] = make_float2(v[r + 8].x * wb.x, v[r + 8].y * wb.y);
    }
  }
  if (tile_pending) tile_ready();   // warps without frames still take part in the barrier
  if (fa < fb) {   // this run's open half-frame, for the next warp
#pragma unroll
    for (int r = 0; r < 8; ++r) Y1[32 * r + lane] = tail[r];
  }
  __syncthreads();
  if (fa < fb && fa > F0) {
    const float2* prev = Y1 - 2 * kYP;   // the previous warp's buffer (its run is not empty when this one is not)
    float2* dst = reinterpret_cast<float2*>(ob + (int64_t)(fa - 1) * kHop) + lane;
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      const float2 wa = s_ws[32 * r + lane], wb = s_ws[32 * (r + 8) + lane];
      const float2 cur = dst[32 * r], tl = prev[32 * r + lane];
      const float nx = 4.f * fmaf(wa.x, wa.x, wb.x * wb.x), ny = 4.f * fmaf(wa.y, wa.y, wb.y * wb.y);
      float2 o;
      o.x = (tl.x + cur.x) / (nx > 1e-10f ? nx : 1.f);
      o.y = (tl.y + cur.y) / (ny > 1e-10f ? ny : 1.f);
      dst[32 * r] = o;
      my_peak = fmaxf(my_peak, fmaxf(fabsf(o.x), fabsf(o.y)));
    }
  }
  if (peak != nullptr) {
    my_peak = warp_max(my_peak);
    if (lane == 0) s_peak[warp] = my_peak;
    __syncthreads();
    if (threadIdx.x == 0) {
      float m = 0.f;
#pragma unroll
      for (int i = 0; i < kWarps; ++i) m = fmaxf(m, s_peak[i]);
      atomicMax(reinterpret_cast<unsigned int*>(peak + b), __float_as_uint(m));
    }
  }
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
static int units_per_cta(int B, int n, int sms, int cap) {
  int per_utt = (16 * sms + B - 1) / B;   // aim at >= 16 CTAs per SM over the launch
  if (per_utt < 1) per_utt = 1;
  int u = (n + per_utt - 1) / per_utt;
  if (u < 4 * kWarps) u = 4 * kWarps;
  if (u > cap) u = cap;
  if (u > n) u = n;
  return u < 1 ? 1 : u;
}

int cov_chunks1024(int B, int T) {
  const int fpc = units_per_cta(B, T, num_sms(), kTileFrames);
  return (T + fpc - 1) / fpc;
}

static constexpr size_t kSmemFrames = (512 + (size_t)kWarps * f512::kSmemComplex + (size_t)kWarps * 2 * kYP) * sizeof(float2);
static constexpr size_t kSmemTile = (size_t)kF * kTilePitch * sizeof(float);
static constexpr size_t kSmemCov = kSmemFrames + kSmemTile;
static constexpr size_t kSmemApply = kSmemFrames + (512 + 2 * (size_t)kYP) * sizeof(float2) + kSmemTile;
static constexpr size_t kSmemFeat = (512 + (size_t)kWarps * f512::kSmemComplex + (size_t)kWarps * 2 * kYP) * sizeof(float2) +
                                    2 * (size_t)kF * kFeatPitch * sizeof(float);
static_assert(kSmemFrames - 512 * sizeof(float2) >= (size_t)kWarps * 5 * kFP * sizeof(float), "reduction staging must fit");

static int tables2(Tables* t512, Tables* t1024) {
  int rc = tables_for(512, t512);
  if (rc) return rc;
  return tables_for(kN, t1024);
}

// plain batch [B][2][L], or windows of planar recordings [R][2][rec_len] read in place (B = R * n_windows)
static WaveView view_of(const AvzChunkView* cv, int64_t L) {
  if (cv == nullptr) return WaveView{2 * L, L, 1, 0, (int)L};
  return WaveView{2 * cv->rec_len, cv->rec_len, cv->n_windows, cv->stride, (int)cv->rec_len};
}

int launch_features(const float* mix, int B, int64_t L, int mode, float* X, cudaStream_t st, const AvzChunkView* cv) {
  if (L >= (1ll << 31) - 2 * kN) return set_error(AVZ_EINVAL, "L too large");
  if (B > 65535) return set_error(AVZ_EINVAL, "B > 65535");
  Tables t5, t10;
  int rc = tables2(&t5, &t10);
  if (rc) return rc;
  const int T = (int)avz_num_frames(L, kN, kHop);
  AVZ_CUDA_OK(cudaFuncSetAttribute(k1024_features<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemFeat));
  AVZ_CUDA_OK(cudaFuncSetAttribute(k1024_features<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemFeat));
  const int64_t n_tiles = (int64_t)B * ((T + kFeatFrames - 1) / kFeatFrames);
  const int64_t resident = 2 * (int64_t)num_sms();
  const unsigned grid = (unsigned)(n_tiles < resident ? n_tiles : resident);
  if (mode == AVZ_FEAT_PHYSICS_NHWC)
    k1024_features<true><<<grid, kWarps * 32, kSmemFeat, st>>>(mix, view_of(cv, L), (int)L, T, B, mode, X, t5, t10);
  else
    k1024_features<false><<<grid, kWarps * 32, kSmemFeat, st>>>(mix, view_of(cv, L), (int)L, T, B, mode, X, t5, t10);
  AVZ_LAUNCH_OK("k1024_features");
  return AVZ_OK;
}

int64_t spec_ws_bytes1024(int B, int T) { return (int64_t)B * T * 2 * kYP * (int64_t)sizeof(float2); }

int launch_mask_cov(const float* mix, const float* mask, int B, int64_t L, float sqrt_eps, float* part, int* chunks_out,
                    void* spec, cudaStream_t st, const AvzChunkView* cv) {
  if (L >= (1ll << 31) - 2 * kN) return set_error(AVZ_EINVAL, "L too large");
  if (B > 65535) return set_error(AVZ_EINVAL, "B > 65535");
  Tables t5, t10;
  int rc = tables2(&t5, &t10);
  if (rc) return rc;
  const int T = (int)avz_num_frames(L, kN, kHop);
  const int fpc = units_per_cta(B, T, num_sms(), kTileFrames);
  const int chunks = (T + fpc - 1) / fpc;
  *chunks_out = chunks;
  AVZ_CUDA_OK(cudaFuncSetAttribute(k1024_cov<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemCov));
  AVZ_CUDA_OK(cudaFuncSetAttribute(k1024_cov<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemCov));
  prof_begin(PROF_COV, st);
  if (spec)
    k1024_cov<true><<<dim3(chunks, B), kWarps * 32, kSmemCov, st>>>(mix, view_of(cv, L), mask, (int)L, T, fpc, sqrt_eps,
                                                                    part, (float2*)spec, t5, t10);
  else
    k1024_cov<false><<<dim3(chunks, B), kWarps * 32, kSmemCov, st>>>(mix, view_of(cv, L), mask, (int)L, T, fpc, sqrt_eps,
                                                                     part, nullptr, t5, t10);
  prof_end(PROF_COV, st);
  AVZ_LAUNCH_OK("k1024_cov");
  return AVZ_OK;
}

int launch_apply(const float* mix, const void* spec, const float* w, const float* mask, int gain_mode, float post_floor,
                 int B, int64_t L, float* out, float* peak, cudaStream_t st, const AvzChunkView* cv) {
  if (L >= (1ll << 31) - 2 * kN) return set_error(AVZ_EINVAL, "L too large");
  if (B > 65535) return set_error(AVZ_EINVAL, "B > 65535");
  Tables t5, t10;
  int rc = tables2(&t5, &t10);
  if (rc) return rc;
  const int T = (int)avz_num_frames(L, kN, kHop);
  if (T < 2) return AVZ_OK;
  // frames per CTA = blocks + 1 (one warm-up frame): keep it a multiple of the warp count so the runs are even
  int bpc = units_per_cta(B, T - 1, num_sms(), kTileFrames - 1);
  bpc = ((bpc + 1 + kWarps - 1) / kWarps) * kWarps - 1;   // <= kTileFrames - 1: the CTA's mask tile covers its frames
  if (bpc > T - 1) bpc = T - 1;
  const int chunks = (T - 1 + bpc - 1) / bpc;
  AVZ_CUDA_OK(cudaFuncSetAttribute(k1024_apply<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemApply));
  AVZ_CUDA_OK(cudaFuncSetAttribute(k1024_apply<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemApply));
  prof_begin(PROF_APPLY, st);
  if (spec)
    k1024_apply<true><<<dim3(chunks, B), kWarps * 32, kSmemApply, st>>>(
        nullptr, view_of(cv, L), reinterpret_cast<const float2*>(spec), reinterpret_cast<const float2*>(w), mask, gain_mode,
        post_floor, (int)L, T, bpc, out, peak, t5, t10);
  else
    k1024_apply<false><<<dim3(chunks, B), kWarps * 32, kSmemApply, st>>>(
        mix, view_of(cv, L), nullptr, reinterpret_cast<const float2*>(w), mask, gain_mode, post_floor, (int)L, T, bpc, out,
        peak, t5, t10);
  prof_end(PROF_APPLY, st);
  AVZ_LAUNCH_OK("k1024_apply");
  return AVZ_OK;
}

}  // namespace o1024
}  // namespace avz

// Generic-size (n_fft in {256, 512, 1024}) fused kernels: one warp per frame, shared-memory FFT.
//
//   k_stft         frame + window + FFT (two real channels per complex transform) -> spectrum
//   k_cov          pass A: spectra of (tgt, int) -> IBM bits (exact-tie path in float64), spectra of the
//                  two mics -> mask-weighted 2x2 Hermitian covariance partial sums (no spectrum stored)
//   k_synth        pass B / iSTFT: [STFT(mix) -> w^H y -> post-filter | given spectrum] -> inverse FFT ->
//                  window -> deterministic overlap-add -> / sum w^2 -> trimmed output (+ running peak)
//
// Reference blocks replaced: rt_av_zoom/core/oracle_debug.py:42-94 (see include/avzoom.h per entry).
#include "avz_common.cuh"

#include <cstdlib>

namespace avz {

// n_fft = 512 fast path (avz_opt512.cu)
namespace o512 {
int cov_chunks512(int B, int T);
int64_t ws_bytes512(int B, int T);
int64_t spec_ws_bytes512(int B, int T);
int launch_stream_step(float* state, const float* hop_in, const float* noise_w, const float* dvec, int n_streams,
                       int t, int t_end, float lam, const AvzMvdrCfg* cfg, float* hop_out, cudaStream_t st);
template <int HOP>
int launch_features(const float* mix, int B, int64_t L, int wrapped, float* X, cudaStream_t st);
int launch_ibm_exact(const float* tgt, const float* itf, int B, int64_t L, int hop, uint32_t* ibm_bits, void* ws16,
                     cudaStream_t st);
template <int HOP>
int launch_ibm_cov(const float* mix, const float* tgt, const float* itf, const float* mask, int B, int64_t L,
                   float sqrt_eps, uint32_t* ibm_bits, float* part, int* chunks_out, void* spec, cudaStream_t st,
                   const CovTailArgs* tail = nullptr, int sparse = 0);
int64_t fused_ws_bytes(int B, int64_t L, int hop);
template <int HOP>
int launch_oracle_fused(const float* mix, const float* tgt, const float* itf, int B, int64_t L, const AvzMvdrCfg* cfg,
                        float norm_eps, float peak_eps, const float* dvec, uint32_t* ibm_bits, float* R, float* msum,
                        float* w, float* out, float* peak, void* ws, cudaStream_t st);
template <int HOP>
int launch_apply(const float* mix, const void* spec, const float* w, const uint32_t* ibm_bits, const float* mask,
                 int gain_mode, float post_floor, int B, int64_t L, float* out, float* peak, int fuse_norm,
                 float peak_eps, int mask_staged, cudaStream_t st, int sparse = 0);
}  // namespace o512

// n_fft = 1024 / hop 512 fast path (avz_opt1024.cu)
namespace o1024 {
int cov_chunks1024(int B, int T);
int launch_features(const float* mix, int B, int64_t L, int mode, float* X, cudaStream_t st, const AvzChunkView* cv = nullptr);
int64_t spec_ws_bytes1024(int B, int T);
int launch_mask_cov(const float* mix, const float* mask, int B, int64_t L, float sqrt_eps, float* part, int* chunks_out,
                    void* spec, cudaStream_t st, const AvzChunkView* cv = nullptr);
int launch_apply(const float* mix, const void* spec, const float* w, const float* mask, int gain_mode, float post_floor,
                 int B, int64_t L, float* out, float* peak, cudaStream_t st, const AvzChunkView* cv = nullptr);
}  // namespace o1024

// The register-resident 512-point path serves n_fft 512 with hop 128 / 256.  Experiment builds (-DAVZ_EXPERIMENT,
// tools/build_exp.sh) can force the generic kernels with AVZ_FORCE_GENERIC=1 for A/B checks of the two implementations
// against each other; the release library reads no environment variable - its results depend on its arguments only.
static bool force_generic() {
#ifdef AVZ_EXPERIMENT
  static const bool forced = [] {
    const char* e = getenv("AVZ_FORCE_GENERIC");
    return e && e[0] == '1';
  }();
  return forced;
#else
  return false;
#endif
}
static bool use_opt512(int n_fft, int hop) { return !force_generic() && n_fft == 512 && (hop == 128 || hop == 256); }
// ... and n_fft 1024 / hop 512 (the learned pipelines' shape) on the same 512-point transform (even/odd split).
static bool use_opt1024(int n_fft, int hop) { return !force_generic() && n_fft == 1024 && hop == 512; }

template <int N>
struct Geo {
  static constexpr int F = N / 2 + 1;
  static constexpr int FW = N / 64 + 1;                 // 32-bit words per frame of IBM bits
  static constexpr int BINS_PER_LANE = N / 64 + 1;      // bins k = lane + 32 i, i < BINS_PER_LANE, k <= N/2
  static constexpr int FP = (F + 31) / 32 * 32;         // padded bin count for partial sums
  static constexpr int WARPS = (N <= 512) ? 8 : 4;
};

__host__ __device__ inline int cov_chunks(int B, int T, int warps, int sms) {
  // enough (utterance, frame-chunk) blocks to fill the machine twice; whole utterances when B is large
  int chunks = 1;
  if (B < 2 * sms) chunks = (2 * sms + B - 1) / B;
  int max_chunks = T / (2 * warps);
  if (max_chunks < 1) max_chunks = 1;
  if (chunks > max_chunks) chunks = max_chunks;
  return chunks;
}

// ------------------------------------------------------------------------------------------
// STFT
// ------------------------------------------------------------------------------------------
template <int N>
__global__ void __launch_bounds__(Geo<N>::WARPS * 32)
k_stft(const float* __restrict__ x, int C, int64_t L, int T, int hop, float2* __restrict__ Y, Tables tb,
       int frames_per_block) {
  constexpr int WARPS = Geo<N>::WARPS;
  constexpr int F = Geo<N>::F;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float2* s_tw = reinterpret_cast<float2*>(smem_raw);
  float* s_win = reinterpret_cast<float*>(s_tw + N);
  float2* bufs = reinterpret_cast<float2*>(s_win + N);
  for (int i = threadIdx.x; i < N; i += WARPS * 32) {
    s_tw[i] = tb.tw[i];
    s_win[i] = tb.win[i];
  }
  __syncthreads();
  const int warp = warp_id_uniform(), lane = threadIdx.x & 31;
  const int b = blockIdx.z, c0 = 2 * blockIdx.y;
  const bool has_b = (c0 + 1) < C;
  const float* sa = x + ((int64_t)b * C + c0) * L;
  const float* sb = has_b ? sa + L : nullptr;
  float2* b0 = bufs + (size_t)warp * 2 * N;
  float2* b1 = b0 + N;
  float2* Ya = Y + ((int64_t)b * C + c0) * F * (int64_t)T;
  float2* Yb = Ya + (int64_t)F * T;
  const int t_begin = blockIdx.x * frames_per_block;
  const int t_end = min(T, t_begin + frames_per_block);
  for (int t = t_begin + warp; t < t_end; t += WARPS) {
    load_frame_pair<N>(b0, sa, sb, L, (int64_t)t * hop - N / 2, s_win, 2.0f / N, lane);
    __syncwarp();
    const float2* z = warp_fft_smem<N, false>(b0, b1, s_tw, lane);
    for (int k = lane; k <= N / 2; k += kWarp) {
      float2 A, Bv;
      unpack_pair(z[k], z[(N - k) & (N - 1)], A, Bv);
      Ya[(int64_t)k * T + t] = A;
      if (has_b) Yb[(int64_t)k * T + t] = Bv;
    }
    __syncwarp();
  }
}

// Features straight from the waveform: STFT of the two mics fused with log-magnitude + IPD
// (full_audio.../inference.py:90-94); the spectrum is never written.
template <int N>
__global__ void __launch_bounds__(Geo<N>::WARPS * 32)
k_wave_features(const float* __restrict__ mix, int64_t L, int T, int hop, int mode, float* __restrict__ X, Tables tb,
                int frames_per_block) {
  constexpr int WARPS = Geo<N>::WARPS;
  constexpr int F = Geo<N>::F;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float2* s_tw = reinterpret_cast<float2*>(smem_raw);
  float* s_win = reinterpret_cast<float*>(s_tw + N);
  float2* bufs = reinterpret_cast<float2*>(s_win + N);
  for (int i = threadIdx.x; i < N; i += WARPS * 32) {
    s_tw[i] = tb.tw[i];
    s_win[i] = tb.win[i];
  }
  __syncthreads();
  const int warp = warp_id_uniform(), lane = threadIdx.x & 31;
  const int b = blockIdx.y;
  const float* m0 = mix + (int64_t)b * 2 * L;
  const float* m1 = m0 + L;
  float2* b0 = bufs + (size_t)warp * 2 * N;
  float2* b1 = b0 + N;
  const int t_begin = blockIdx.x * frames_per_block;
  const int t_end = min(T, t_begin + frames_per_block);
  for (int t = t_begin + warp; t < t_end; t += WARPS) {
    load_frame_pair<N>(b0, m0, m1, L, (int64_t)t * hop - N / 2, s_win, 2.0f / N, lane);
    __syncwarp();
    const float2* z = warp_fft_smem<N, false>(b0, b1, s_tw, lane);
    for (int k = lane; k <= N / 2; k += kWarp) {
      float2 Y0, Y1;
      unpack_pair(z[k], z[(N - k) & (N - 1)], Y0, Y1);
      float lm, ipd;
      feature_values(Y0, Y1, lm, ipd);
      store_features(X, mode, b, k, t, F, T, lm, ipd);
    }
    __syncwarp();
  }
}

// ------------------------------------------------------------------------------------------
// pass A: IBM + masked covariance partial sums
// ------------------------------------------------------------------------------------------
// Exact decision |S_int[k]| > |S_tgt[k]| for one frame in float64 (direct DFT, whole warp cooperates).
template <int N>
__device__ __forceinline__ bool ibm_exact(const float* __restrict__ tgt, const float* __restrict__ itf, int64_t L,
                                          int64_t start, int k, const Tables& tb, int lane) {
  double tr = 0, ti = 0, ir = 0, ii = 0;
  for (int n = lane; n < N; n += kWarp) {
    const int64_t i = start + n;
    if (i >= 0 && i < L) {
      const double w = tb.win_d[n];
      const double2 e = tb.tw_d[(n * k) & (N - 1)];
      const double a = w * (double)__ldg(tgt + i);
      const double c = w * (double)__ldg(itf + i);
      tr = fma(a, e.x, tr);
      ti = fma(a, e.y, ti);
      ir = fma(c, e.x, ir);
      ii = fma(c, e.y, ii);
    }
  }
  tr = warp_sum(tr);
  ti = warp_sum(ti);
  ir = warp_sum(ir);
  ii = warp_sum(ii);
  return (ir * ir + ii * ii) > (tr * tr + ti * ti);
}

enum { COV_IBM = 0, COV_MASK = 1 };

// part layout: [B][chunks][5][FP] = (R00, R11, Re R01, Im R01, sum m), un-normalised.
template <int N, int MODE>
__global__ void __launch_bounds__(Geo<N>::WARPS * 32)
k_cov(const float* __restrict__ mix, const float* __restrict__ tgt, const float* __restrict__ itf,
      const float* __restrict__ mask, int64_t L, int T, int hop, int frames_per_chunk, float sqrt_eps,
      uint32_t* __restrict__ ibm_bits, float* __restrict__ part, Tables tb) {
  constexpr int WARPS = Geo<N>::WARPS;
  constexpr int F = Geo<N>::F;
  constexpr int FW = Geo<N>::FW;
  constexpr int BPL = Geo<N>::BINS_PER_LANE;
  constexpr int FP = Geo<N>::FP;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float2* s_tw = reinterpret_cast<float2*>(smem_raw);
  float* s_win = reinterpret_cast<float*>(s_tw + N);
  float2* bufs = reinterpret_cast<float2*>(s_win + N);
  float* s_acc = reinterpret_cast<float*>(bufs + (size_t)WARPS * 2 * N);  // [WARPS][5][FP]
  for (int i = threadIdx.x; i < N; i += WARPS * 32) {
    s_tw[i] = tb.tw[i];
    s_win[i] = tb.win[i];
  }
  for (int i = threadIdx.x; i < WARPS * 5 * FP; i += WARPS * 32) s_acc[i] = 0.f;
  __syncthreads();
  const int warp = warp_id_uniform(), lane = threadIdx.x & 31;
  const int b = blockIdx.y, chunk = blockIdx.x, chunks = gridDim.x;
  const float* m0 = mix + (int64_t)b * 2 * L;
  const float* m1 = m0 + L;
  const float* tg = (MODE == COV_IBM) ? tgt + (int64_t)b * L : nullptr;
  const float* it = (MODE == COV_IBM) ? itf + (int64_t)b * L : nullptr;
  float2* b0 = bufs + (size_t)warp * 2 * N;
  float2* b1 = b0 + N;
  float* acc = s_acc + (size_t)warp * 5 * FP;
  const int t_begin = chunk * frames_per_chunk;
  const int t_end = min(T, t_begin + frames_per_chunk);

  for (int t = t_begin + warp; t < t_end; t += WARPS) {
    const int64_t start = (int64_t)t * hop - N / 2;
    float wgt[BPL];  // noise weight of this lane's bins in this frame
    if (MODE == COV_IBM) {
      float e2 = load_frame_pair<N>(b0, tg, it, L, start, s_win, 2.0f / N, lane);
      e2 = warp_sum(e2);
      __syncwarp();
      const float2* z = warp_fft_smem<N, false>(b0, b1, s_tw, lane);
      const float delta = 4e-6f * sqrtf(e2);  // bound on the float32 FFT error of one bin
      uint32_t* bits_t = ibm_bits + ((int64_t)b * T + t) * FW;
#pragma unroll
      for (int i = 0; i < BPL; ++i) {
        const int k = lane + 32 * i;
        bool bit = false, amb = false;
        if (k <= N / 2) {
          float2 St, Si;
          unpack_pair(z[k], z[(N - k) & (N - 1)], St, Si);
          const float pt = cabs2(St), pi = cabs2(Si);
          bit = pi > pt;
          const float mx = fmaxf(pt, pi);
          // too close to call in float32 (but not the exact 0 == 0 tie of silence): redo in float64
          amb = (e2 > 0.f) && (fabsf(pi - pt) <= 4.f * delta * sqrtf(mx) + 2.f * delta * delta);
        }
        unsigned am = __ballot_sync(kFull, amb);
        while (am) {
          const int src = __ffs(am) - 1;
          const bool r = ibm_exact<N>(tg, it, L, start, src + 32 * i, tb, lane);
          if (lane == src) bit = r;
          am &= am - 1;
        }
        const unsigned word = __ballot_sync(kFull, bit);
        if (lane == 0) bits_t[i] = word;
        wgt[i] = bit ? 1.f : 0.f;
      }
      __syncwarp();
    } else {
#pragma unroll
      for (int i = 0; i < BPL; ++i) {
        const int k = lane + 32 * i;
        wgt[i] = (k <= N / 2) ? 1.f - __ldg(mask + ((int64_t)b * F + k) * T + t) : 0.f;
      }
    }
    load_frame_pair<N>(b0, m0, m1, L, start, s_win, 2.0f / N, lane);
    __syncwarp();
    const float2* z = warp_fft_smem<N, false>(b0, b1, s_tw, lane);
#pragma unroll
    for (int i = 0; i < BPL; ++i) {
      const int k = lane + 32 * i;
      if (k <= N / 2) {
        float2 Y0, Y1;
        unpack_pair(z[k], z[(N - k) & (N - 1)], Y0, Y1);
        const float m = wgt[i];
        const float ms = (MODE == COV_MASK) ? m + sqrt_eps : m;
        const float2 c01 = cmulc(Y0, Y1);
        acc[0 * FP + k] = fmaf(ms, cabs2(Y0), acc[0 * FP + k]);
        acc[1 * FP + k] = fmaf(ms, cabs2(Y1), acc[1 * FP + k]);
        acc[2 * FP + k] = fmaf(ms, c01.x, acc[2 * FP + k]);
        acc[3 * FP + k] = fmaf(ms, c01.y, acc[3 * FP + k]);
        acc[4 * FP + k] += m;
      }
    }
    __syncwarp();
  }
  __syncthreads();
  float* dst = part + ((int64_t)b * chunks + chunk) * 5 * FP;
  for (int i = threadIdx.x; i < 5 * FP; i += WARPS * 32) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < WARPS; ++w) s += s_acc[(size_t)w * 5 * FP + i];  // fixed order: bit-stable reruns
    dst[i] = s;
  }
}

// Sum the chunk partials (float64), normalise: R = sum / (sum m + norm_eps).  The kernel is a pure load-latency chain
// per thread, so the chunks of a bin are split over kFinSlices threads (slice s takes chunks s, s + kFinSlices, ...)
// whose float64 sums meet in shared memory in a fixed order: reruns stay bit-identical.
constexpr int kFinSlices = 4;
constexpr int kFinBins = 64;
__global__ void __launch_bounds__(kFinBins * kFinSlices)
k_cov_finalize(const float* __restrict__ part, int B, int F, int FP, int chunks, float norm_eps,
               float4* __restrict__ R, float* __restrict__ msum) {
  __shared__ double s_sum[kFinSlices][5][kFinBins];
  const int lane64 = threadIdx.x & (kFinBins - 1), slice = threadIdx.x / kFinBins;
  const int idx = blockIdx.x * kFinBins + lane64;
  const bool ok = idx < B * F;
  const int b = ok ? idx / F : 0, k = ok ? idx - b * F : 0;
  double s[5] = {0, 0, 0, 0, 0};
  if (ok) {
#pragma unroll 4   // independent loads: keep several chunks in flight
    for (int c = slice; c < chunks; c += kFinSlices) {
      const float* p = part + ((int64_t)b * chunks + c) * 5 * FP;
#pragma unroll
      for (int j = 0; j < 5; ++j) s[j] += (double)p[j * FP + k];
    }
  }
#pragma unroll
  for (int j = 0; j < 5; ++j) s_sum[slice][j][lane64] = s[j];
  __syncthreads();
  if (slice != 0 || !ok) return;
#pragma unroll
  for (int j = 0; j < 5; ++j) {
    double t = s_sum[0][j][lane64];
#pragma unroll
    for (int q = 1; q < kFinSlices; ++q) t += s_sum[q][j][lane64];
    s[j] = t;
  }
  const double inv = 1.0 / (s[4] + (double)norm_eps);
  R[idx] = make_float4((float)(s[0] * inv), (float)(s[1] * inv), (float)(s[2] * inv), (float)(s[3] * inv));
  msum[idx] = (float)s[4];
}

// ------------------------------------------------------------------------------------------
// pass B / iSTFT: synthesis with deterministic overlap-add
// ------------------------------------------------------------------------------------------
enum { SRC_SPEC = 0, SRC_MIX = 1 };
enum { GAIN_NONE = 0, GAIN_BITS = 1, GAIN_FLOOR = 2, GAIN_MASK = 3 };

template <int N, int SRC>
__global__ void __launch_bounds__(Geo<N>::WARPS * 32)
k_synth(const float* __restrict__ mix, const float2* __restrict__ spec, const float2* __restrict__ w,
        const uint32_t* __restrict__ ibm_bits, const float* __restrict__ mask, int gain_mode, float post_floor,
        int64_t L, int T, int hop, int blocks_per_chunk, float* __restrict__ out, float* __restrict__ peak,
        Tables tb) {
  constexpr int WARPS = Geo<N>::WARPS;
  constexpr int F = Geo<N>::F;
  constexpr int FW = Geo<N>::FW;
  constexpr int RING = WARPS + 8;  // >= WARPS + R - 1 for every supported overlap factor R <= 8
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float2* s_tw = reinterpret_cast<float2*>(smem_raw);
  float* s_win = reinterpret_cast<float*>(s_tw + N);
  float2* bufs = reinterpret_cast<float2*>(s_win + N);
  float* s_ring = reinterpret_cast<float*>(bufs + (size_t)WARPS * 2 * N);  // [RING][N] windowed frames
  float2* s_a = reinterpret_cast<float2*>(s_ring + (size_t)RING * N);      // [F] beamform coefficients
  float2* s_b = s_a + F;
  __shared__ float s_peak[WARPS];

  const int R = N / hop;
  const int b = blockIdx.y;
  for (int i = threadIdx.x; i < N; i += WARPS * 32) {
    s_tw[i] = tb.tw[i];
    s_win[i] = tb.win[i];
  }
  if (SRC == SRC_MIX) {
    // S[k] = conj(w0) Y0 + conj(w1) Y1 with Y0 = (Z[k] + conj Z[N-k])/2, Y1 = -i (Z[k] - conj Z[N-k])/2
    //      = a[k] Z[k] + b[k] conj(Z[N-k]),  a = (conj w0 - i conj w1)/2,  b = (conj w0 + i conj w1)/2
    for (int k = threadIdx.x; k < F; k += WARPS * 32) {
      const float2 w0 = w[((int64_t)b * F + k) * 2 + 0];
      const float2 w1 = w[((int64_t)b * F + k) * 2 + 1];
      // conj w0 = (w0.x, -w0.y);  i conj w1 = (w1.y, w1.x)
      s_a[k] = make_float2(0.5f * (w0.x - w1.y), 0.5f * (-w0.y - w1.x));
      s_b[k] = make_float2(0.5f * (w0.x + w1.y), 0.5f * (-w0.y + w1.x));
    }
  }
  __syncthreads();

  const int warp = warp_id_uniform(), lane = threadIdx.x & 31;
  float2* b0 = bufs + (size_t)warp * 2 * N;
  float2* b1 = b0 + N;
  const float* m0 = (SRC == SRC_MIX) ? mix + (int64_t)b * 2 * L : nullptr;
  const float* m1 = (SRC == SRC_MIX) ? m0 + L : nullptr;
  const float2* Sb = (SRC == SRC_SPEC) ? spec + (int64_t)b * F * (int64_t)T : nullptr;

  // Output hop-blocks in extended (untrimmed) coordinates: block g covers [g*hop, (g+1)*hop); the trimmed
  // output is blocks [R/2, R/2 + T - 1).  Block g sums frames t in [g-R+1, g] intersected with [0, T-1].
  const int g_lo = R / 2, g_hi = R / 2 + T - 1;
  const int g0 = g_lo + blockIdx.x * blocks_per_chunk;
  const int g1 = min(g_hi, g0 + blocks_per_chunk);
  const int64_t out_len = (int64_t)(T - 1) * hop;
  float* ob = out + (int64_t)b * out_len;
  float my_peak = 0.f;
  const int t_first = max(0, g0 - R + 1);

  for (int tile = t_first; tile < g1; tile += WARPS) {
    const int t = tile + warp;
    if (t < T && t < g1) {
      float2* g;  // spectrum to invert, Hermitian-filled
      if (SRC == SRC_MIX) {
        load_frame_pair<N>(b0, m0, m1, L, (int64_t)t * hop - N / 2, s_win, 2.0f / N, lane);
        __syncwarp();
        float2* z = warp_fft_smem<N, false>(b0, b1, s_tw, lane);
        g = (z == b0) ? b1 : b0;
        for (int k = lane; k <= N / 2; k += kWarp) {
          const float2 zk = z[k], zm = z[(N - k) & (N - 1)];
          float2 s = cadd(cmul(s_a[k], zk), cmulc(s_b[k], zm));
          float gain = 1.f;
          if (gain_mode == GAIN_BITS) {
            const uint32_t word = __ldg(ibm_bits + ((int64_t)b * T + t) * FW + (k >> 5));
            gain = ((word >> (k & 31)) & 1u) ? 0.f : 1.f;  // 1 - noise mask
          } else if (gain_mode == GAIN_FLOOR) {
            gain = fmaxf(__ldg(mask + ((int64_t)b * F + k) * T + t), post_floor);
          } else if (gain_mode == GAIN_MASK) {
            gain = __ldg(mask + ((int64_t)b * F + k) * T + t);
          }
          s.x *= gain;
          s.y *= gain;
          if (k == 0 || k == N / 2) {
            g[k] = make_float2(s.x, 0.f);  // c2r transforms ignore Im(DC), Im(Nyquist)
          } else {
            g[k] = s;
            g[N - k] = make_float2(s.x, -s.y);
          }
        }
        __syncwarp();
        float2* other = (g == b0) ? b1 : b0;
        const float2* xr = warp_fft_smem<N, true>(g, other, s_tw, lane);
        float* slot = s_ring + (size_t)(t % RING) * N;
        for (int n = lane; n < N; n += kWarp) slot[n] = xr[n].x * 0.5f * s_win[n];  // irfft * sum(w) = 0.5 * sum
      } else {
        for (int k = lane; k <= N / 2; k += kWarp) {
          const float2 s = __ldg(Sb + (int64_t)k * T + t);
          if (k == 0 || k == N / 2) {
            b0[k] = make_float2(s.x, 0.f);
          } else {
            b0[k] = s;
            b0[N - k] = make_float2(s.x, -s.y);
          }
        }
        __syncwarp();
        const float2* xr = warp_fft_smem<N, true>(b0, b1, s_tw, lane);
        float* slot = s_ring + (size_t)(t % RING) * N;
        for (int n = lane; n < N; n += kWarp) slot[n] = xr[n].x * 0.5f * s_win[n];
      }
    }
    __syncthreads();
    // emit the blocks completed by this tile
    const int e0 = max(g0, tile);
    const int e1 = min(g1, tile + WARPS);
    for (int idx = threadIdx.x; idx < (e1 - e0) * hop; idx += WARPS * 32) {
      const int gi = idx / hop, p = idx - gi * hop;
      const int gblk = e0 + gi;
      float acc = 0.f, nrm = 0.f;
      for (int q = R - 1; q >= 0; --q) {  // ascending frame index, like the reference's loop
        const int tq = gblk - q;
        if (tq >= 0 && tq < T) {
          const float wv = s_win[q * hop + p];
          acc += s_ring[(size_t)(tq % RING) * N + q * hop + p];
          nrm = fmaf(wv, wv, nrm);
        }
      }
      const float v = acc / (nrm > 1e-10f ? nrm : 1.0f);
      ob[(int64_t)(gblk - g_lo) * hop + p] = v;
      my_peak = fmaxf(my_peak, fabsf(v));
    }
    __syncthreads();
  }
  if (peak != nullptr) {
    my_peak = warp_max(my_peak);
    if (lane == 0) s_peak[warp] = my_peak;
    __syncthreads();
    if (threadIdx.x == 0) {
      float m = 0.f;
      for (int i = 0; i < WARPS; ++i) m = fmaxf(m, s_peak[i]);
      // non-negative floats order like their bit patterns; max is order-independent -> deterministic
      atomicMax(reinterpret_cast<unsigned int*>(peak + b), __float_as_uint(m));
    }
  }
}

__global__ void k_peak_normalise(float* __restrict__ x, int64_t n, const float* __restrict__ peak, float peak_eps) {
  const int b = blockIdx.y;
  const float den = peak[b] + peak_eps;  // true division, like `s_out /= np.max(np.abs(s_out))`: the peak maps to 1.0
  float* xb = x + (int64_t)b * n;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if ((n & 3) == 0 && (reinterpret_cast<uintptr_t>(xb) & 15) == 0) {
    // 16-byte accesses, four independent loads in flight per thread (the kernel is pure HBM streaming)
    float4* x4 = reinterpret_cast<float4*>(xb);
    const int64_t n4 = n >> 2;
    int64_t i = tid;
    for (; i + 3 * stride < n4; i += 4 * stride) {
      float4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) v[u] = x4[i + u * stride];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        v[u] = make_float4(__fdiv_rn(v[u].x, den), __fdiv_rn(v[u].y, den), __fdiv_rn(v[u].z, den), __fdiv_rn(v[u].w, den));
        x4[i + u * stride] = v[u];
      }
    }
    for (; i < n4; i += stride) {
      float4 v = x4[i];
      x4[i] = make_float4(__fdiv_rn(v.x, den), __fdiv_rn(v.y, den), __fdiv_rn(v.z, den), __fdiv_rn(v.w, den));
    }
  } else {
    for (int64_t i = tid; i < n; i += stride) xb[i] = __fdiv_rn(xb[i], den);
  }
}

// ------------------------------------------------------------------------------------------
// host launchers
// ------------------------------------------------------------------------------------------
template <int N>
static size_t smem_base() {
  return (size_t)N * sizeof(float2) + (size_t)N * sizeof(float) + (size_t)Geo<N>::WARPS * 2 * N * sizeof(float2);
}

template <int N>
static int launch_stft(const float* x, int B, int C, int64_t L, int hop, float* Y, cudaStream_t st) {
  Tables tb;
  int rc = tables_for(N, &tb);
  if (rc) return rc;
  const int T = (int)avz_num_frames(L, N, hop);
  const size_t smem = smem_base<N>();
  AVZ_CUDA_OK(cudaFuncSetAttribute(k_stft<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int fpb = Geo<N>::WARPS * 4;
  dim3 grid((T + fpb - 1) / fpb, (C + 1) / 2, B);
  k_stft<N><<<grid, Geo<N>::WARPS * 32, smem, st>>>(x, C, L, T, hop, reinterpret_cast<float2*>(Y), tb, fpb);
  AVZ_LAUNCH_OK("k_stft");
  return AVZ_OK;
}

template <int N>
static int launch_wave_features(const float* mix, int B, int64_t L, int hop, int mode, float* X, cudaStream_t st) {
  Tables tb;
  int rc = tables_for(N, &tb);
  if (rc) return rc;
  const int T = (int)avz_num_frames(L, N, hop);
  const size_t smem = smem_base<N>();
  AVZ_CUDA_OK(cudaFuncSetAttribute(k_wave_features<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int fpb = Geo<N>::WARPS * 4;
  dim3 grid((T + fpb - 1) / fpb, B);
  k_wave_features<N><<<grid, Geo<N>::WARPS * 32, smem, st>>>(mix, L, T, hop, mode, X, tb, fpb);
  AVZ_LAUNCH_OK("k_wave_features");
  return AVZ_OK;
}

template <int N, int MODE>
static int launch_cov(const float* mix, const float* tgt, const float* itf, const float* mask, int B, int64_t L,
                      int hop, float sqrt_eps, float norm_eps, uint32_t* ibm_bits, float* R, float* msum, void* ws,
                      cudaStream_t st) {
  Tables tb;
  int rc = tables_for(N, &tb);
  if (rc) return rc;
  const int T = (int)avz_num_frames(L, N, hop);
  const int chunks = cov_chunks(B, T, Geo<N>::WARPS, num_sms());
  const int fpc = (T + chunks - 1) / chunks;
  const size_t smem = smem_base<N>() + (size_t)Geo<N>::WARPS * 5 * Geo<N>::FP * sizeof(float);
  AVZ_CUDA_OK(cudaFuncSetAttribute(k_cov<N, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid(chunks, B);
  k_cov<N, MODE><<<grid, Geo<N>::WARPS * 32, smem, st>>>(mix, tgt, itf, mask, L, T, hop, fpc, sqrt_eps, ibm_bits,
                                                         (float*)ws, tb);
  AVZ_LAUNCH_OK("k_cov");
  const int F = Geo<N>::F;
  k_cov_finalize<<<(B * F + kFinBins - 1) / kFinBins, kFinBins * kFinSlices, 0, st>>>((const float*)ws, B, F, Geo<N>::FP, chunks, norm_eps,
                                                      reinterpret_cast<float4*>(R), msum);
  AVZ_LAUNCH_OK("k_cov_finalize");
  return AVZ_OK;
}

// fast path: IBM + covariance (or mask covariance) for n_fft 512, then the shared finalize kernel
static int launch_cov512(const float* mix, const float* tgt, const float* itf, const float* mask, int B, int64_t L,
                         int hop, float sqrt_eps, float norm_eps, uint32_t* ibm_bits, float* R, float* msum, void* ws,
                         void* spec, cudaStream_t st, int sparse = 0) {
  int chunks = 0;
  int rc = (hop == 128) ? o512::launch_ibm_cov<128>(mix, tgt, itf, mask, B, L, sqrt_eps, ibm_bits, (float*)ws, &chunks, spec,
                                                    st, nullptr, sparse)
                        : o512::launch_ibm_cov<256>(mix, tgt, itf, mask, B, L, sqrt_eps, ibm_bits, (float*)ws, &chunks, spec,
                                                    st, nullptr, sparse);
  if (rc) return rc;
  const int F = 257;
  prof_begin(PROF_FINALIZE, st);
  k_cov_finalize<<<(B * F + kFinBins - 1) / kFinBins, kFinBins * kFinSlices, 0, st>>>((const float*)ws, B, F, Geo<512>::FP, chunks, norm_eps,
                                                      reinterpret_cast<float4*>(R), msum);
  prof_end(PROF_FINALIZE, st);
  AVZ_LAUNCH_OK("k_cov_finalize");
  return AVZ_OK;
}

template <int N, int SRC>
static int launch_synth(const float* mix, const float* spec, const float* w, const uint32_t* ibm_bits,
                        const float* mask, int gain_mode, float post_floor, int B, int64_t L, int T, int hop,
                        float* out, float* peak, cudaStream_t st) {
  Tables tb;
  int rc = tables_for(N, &tb);
  if (rc) return rc;
  constexpr int WARPS = Geo<N>::WARPS;
  const int n_blocks = T - 1;  // output hop-blocks per utterance
  if (n_blocks <= 0) return AVZ_OK;
  int chunks = cov_chunks(B, n_blocks, WARPS, num_sms());
  const int bpc = (n_blocks + chunks - 1) / chunks;
  chunks = (n_blocks + bpc - 1) / bpc;
  const size_t smem = smem_base<N>() + (size_t)(WARPS + 8) * N * sizeof(float) + 2 * (size_t)Geo<N>::F * sizeof(float2);
  AVZ_CUDA_OK(cudaFuncSetAttribute(k_synth<N, SRC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid(chunks, B);
  k_synth<N, SRC><<<grid, WARPS * 32, smem, st>>>(mix, reinterpret_cast<const float2*>(spec),
                                                  reinterpret_cast<const float2*>(w), ibm_bits, mask, gain_mode,
                                                  post_floor, L, T, hop, bpc, out, peak, tb);
  AVZ_LAUNCH_OK("k_synth");
  return AVZ_OK;
}

#define AVZ_DISPATCH_N(n_fft, CALL)                                      \
  switch (n_fft) {                                                       \
    case 256: { constexpr int N_ = 256; return CALL; }                   \
    case 512: { constexpr int N_ = 512; return CALL; }                   \
    case 1024: { constexpr int N_ = 1024; return CALL; }                 \
    default: return avz::set_error(AVZ_EINVAL, "n_fft=%d unsupported", n_fft); \
  }

}  // namespace avz

using namespace avz;

extern "C" {

int avz_stft_f32(const float* x, int B, int C, int64_t L, int n_fft, int hop, float* Y, void* stream) {
  if (!x || !Y || B <= 0 || B > 65535 || C <= 0) return set_error(AVZ_EINVAL, "avz_stft_f32: null pointer or empty batch");
  int rc = check_fft_args(n_fft, hop, L);
  if (rc) return rc;
  AVZ_DISPATCH_N(n_fft, (launch_stft<N_>(x, B, C, L, hop, Y, (cudaStream_t)stream)));
}

int avz_istft_f32(const float* S, int B, int T, int n_fft, int hop, float* x, float* peak, void* stream) {
  if (!S || !x || B <= 0 || B > 65535 || T < 2) return set_error(AVZ_EINVAL, "avz_istft_f32: null pointer, empty batch or T < 2");
  int rc = check_fft_args(n_fft, hop, n_fft);
  if (rc) return rc;
  AVZ_DISPATCH_N(n_fft, (launch_synth<N_, SRC_SPEC>(nullptr, S, nullptr, nullptr, nullptr, GAIN_NONE, 0.f, B, 0, T, hop,
                                                    x, peak, (cudaStream_t)stream)));
}

int avz_peak_normalise_f32(float* x, int B, int64_t n, const float* peak, float peak_eps, void* stream) {
  if (!x || !peak || B <= 0 || B > 65535 || n <= 0) return set_error(AVZ_EINVAL, "avz_peak_normalise_f32: bad argument");
  int gx = (int)((n / 4 + 1023) / 1024);   // ~4 float4 per thread
  if (gx > 64) gx = 64;
  if (gx < 1) gx = 1;
  prof_begin(PROF_NORMALISE, (cudaStream_t)stream);
  k_peak_normalise<<<dim3(gx, B), 256, 0, (cudaStream_t)stream>>>(x, n, peak, peak_eps);
  prof_end(PROF_NORMALISE, (cudaStream_t)stream);
  AVZ_LAUNCH_OK("k_peak_normalise");
  return AVZ_OK;
}

int avz_wave_features_f32(const float* mix, int B, int64_t L, int n_fft, int hop, int mode, float* X, void* stream) {
  if (!mix || !X || B <= 0 || B > 65535 || mode < 0 || mode > 2) return set_error(AVZ_EINVAL, "avz_wave_features_f32: bad argument");
  int rc = check_fft_args(n_fft, hop, L);
  if (rc) return rc;
  if (use_opt512(n_fft, hop) && mode != AVZ_FEAT_PHYSICS_NHWC) {
    const int wrapped = (mode == AVZ_FEAT_LOGMAG_IPD_WRAPPED);
    return (hop == 128) ? o512::launch_features<128>(mix, B, L, wrapped, X, (cudaStream_t)stream)
                        : o512::launch_features<256>(mix, B, L, wrapped, X, (cudaStream_t)stream);
  }
  if (use_opt1024(n_fft, hop)) return o1024::launch_features(mix, B, L, mode, X, (cudaStream_t)stream);
  AVZ_DISPATCH_N(n_fft, (launch_wave_features<N_>(mix, B, L, hop, mode, X, (cudaStream_t)stream)));
}

int64_t avz_ibm_cov_ws_bytes(int B, int64_t L, int n_fft, int hop) {
  if (B <= 0 || check_fft_args(n_fft, hop, L)) return -1;
  const int T = (int)avz_num_frames(L, n_fft, hop);
  const int warps = (n_fft <= 512) ? 8 : 4;
  int chunks = cov_chunks(B, T, warps, num_sms());
  if (n_fft == 512) chunks = max(chunks, o512::cov_chunks512(B, T));
  if (n_fft == 1024) chunks = max(chunks, o1024::cov_chunks1024(B, T));
  const int FP = ((n_fft / 2 + 1) + 31) / 32 * 32;
  int64_t bytes = (int64_t)B * chunks * 5 * FP * (int64_t)sizeof(float);
  if (n_fft == 512) bytes = max(bytes, o512::ws_bytes512(B, T));
  return bytes;
}

int avz_ibm_cov_f32(const float* mix, const float* tgt, const float* itf, int B, int64_t L, int n_fft, int hop,
                    float norm_eps, uint32_t* ibm_bits, float* R, float* msum, void* ws, void* stream) {
  if (!mix || !tgt || !itf || !ibm_bits || !R || !msum || !ws || B <= 0 || B > 65535)
    return set_error(AVZ_EINVAL, "avz_ibm_cov_f32: null pointer or empty batch");
  int rc = check_fft_args(n_fft, hop, L);
  if (rc) return rc;
  if (use_opt512(n_fft, hop))
    return launch_cov512(mix, tgt, itf, nullptr, B, L, hop, 0.f, norm_eps, ibm_bits, R, msum, ws, nullptr,
                         (cudaStream_t)stream);
  AVZ_DISPATCH_N(n_fft, (launch_cov<N_, COV_IBM>(mix, tgt, itf, nullptr, B, L, hop, 0.f, norm_eps, ibm_bits, R, msum,
                                                 ws, (cudaStream_t)stream)));
}

int avz_ibm_exact_f32(const float* tgt, const float* itf, int B, int64_t L, int n_fft, int hop, uint32_t* ibm_bits,
                      void* ws16, void* stream) {
  if (!tgt || !itf || !ibm_bits || !ws16 || B <= 0) return set_error(AVZ_EINVAL, "avz_ibm_exact_f32: bad argument");
  int rc = check_fft_args(n_fft, hop, L);
  if (rc) return rc;
  if (n_fft != 512) return set_error(AVZ_EINVAL, "avz_ibm_exact_f32: n_fft must be 512");
  return o512::launch_ibm_exact(tgt, itf, B, L, hop, ibm_bits, ws16, (cudaStream_t)stream);
}

static int mask_cov1024(const float* mix, const float* mask, int B, int64_t L, float sqrt_eps, float norm_eps, float* R,
                        float* msum, void* ws, void* spec, cudaStream_t st, const AvzChunkView* cv = nullptr) {
  int chunks = 0;
  int rc = o1024::launch_mask_cov(mix, mask, B, L, sqrt_eps, (float*)ws, &chunks, spec, st, cv);
  if (rc) return rc;
  k_cov_finalize<<<(B * 513 + kFinBins - 1) / kFinBins, kFinBins * kFinSlices, 0, st>>>(
      (const float*)ws, B, 513, Geo<1024>::FP, chunks, norm_eps, reinterpret_cast<float4*>(R), msum);
  AVZ_LAUNCH_OK("k_cov_finalize");
  return AVZ_OK;
}

int avz_wave_mask_cov_f32(const float* mix, const float* mask, int B, int64_t L, int n_fft, int hop, float sqrt_eps,
                          float norm_eps, float* R, float* msum, void* ws, void* stream) {
  if (!mix || !mask || !R || !msum || !ws || B <= 0 || B > 65535)
    return set_error(AVZ_EINVAL, "avz_wave_mask_cov_f32: null pointer or empty batch");
  int rc = check_fft_args(n_fft, hop, L);
  if (rc) return rc;
  if (use_opt512(n_fft, hop))
    return launch_cov512(mix, nullptr, nullptr, mask, B, L, hop, sqrt_eps, norm_eps, nullptr, R, msum, ws, nullptr,
                         (cudaStream_t)stream);
  if (use_opt1024(n_fft, hop))
    return mask_cov1024(mix, mask, B, L, sqrt_eps, norm_eps, R, msum, ws, nullptr, (cudaStream_t)stream);
  AVZ_DISPATCH_N(n_fft, (launch_cov<N_, COV_MASK>(mix, nullptr, nullptr, mask, B, L, hop, sqrt_eps, norm_eps, nullptr,
                                                  R, msum, ws, (cudaStream_t)stream)));
}

// ---- "kept spectrum" variants of the fused passes (n_fft 512 fast path only) ---------------------------------
int64_t avz_spec_ws_bytes(int B, int64_t L, int n_fft, int hop) {
  if (B <= 0 || check_fft_args(n_fft, hop, L)) return 0;
  if (use_opt1024(n_fft, hop)) return o1024::spec_ws_bytes1024(B, (int)avz_num_frames(L, n_fft, hop));
  if (!use_opt512(n_fft, hop)) return 0;
  return o512::spec_ws_bytes512(B, (int)avz_num_frames(L, n_fft, hop));
}

int avz_ibm_cov_keep_f32(const float* mix, const float* tgt, const float* itf, int B, int64_t L, int n_fft, int hop,
                         float norm_eps, uint32_t* ibm_bits, float* R, float* msum, void* ws, void* spec, void* stream) {
  if (!mix || !tgt || !itf || !ibm_bits || !R || !msum || !ws || !spec || B <= 0 || B > 65535)
    return set_error(AVZ_EINVAL, "avz_ibm_cov_keep_f32: null pointer or empty batch");
  int rc = check_fft_args(n_fft, hop, L);
  if (rc) return rc;
  if (!use_opt512(n_fft, hop)) return set_error(AVZ_EINVAL, "avz_ibm_cov_keep_f32: n_fft 512, hop 128/256 only");
  return launch_cov512(mix, tgt, itf, nullptr, B, L, hop, 0.f, norm_eps, ibm_bits, R, msum, ws, spec,
                       (cudaStream_t)stream);
}

int avz_ibm_cov_keep_sparse_f32(const float* mix, const float* tgt, const float* itf, int B, int64_t L, int n_fft, int hop,
                                float norm_eps, uint32_t* ibm_bits, float* R, float* msum, void* ws, void* spec, void* stream) {
  if (!mix || !tgt || !itf || !ibm_bits || !R || !msum || !ws || !spec || B <= 0 || B > 65535)
    return set_error(AVZ_EINVAL, "avz_ibm_cov_keep_sparse_f32: null pointer or empty batch");
  int rc = check_fft_args(n_fft, hop, L);
  if (rc) return rc;
  if (!use_opt512(n_fft, hop)) return set_error(AVZ_EINVAL, "avz_ibm_cov_keep_sparse_f32: n_fft 512, hop 128/256 only");
  return launch_cov512(mix, tgt, itf, nullptr, B, L, hop, 0.f, norm_eps, ibm_bits, R, msum, ws, spec,
                       (cudaStream_t)stream, 1);
}

int avz_ibm_cov_keep_postmask_f32(const float* mix, const float* tgt, const float* itf, int B, int64_t L, int n_fft, int hop,
                                  float norm_eps, uint32_t* ibm_bits, float* R, float* msum, void* ws, void* spec,
                                  void* stream) {
  if (!mix || !tgt || !itf || !ibm_bits || !R || !msum || !ws || !spec || B <= 0 || B > 65535)
    return set_error(AVZ_EINVAL, "avz_ibm_cov_keep_postmask_f32: null pointer or empty batch");
  int rc = check_fft_args(n_fft, hop, L);
  if (rc) return rc;
  if (!use_opt512(n_fft, hop)) return set_error(AVZ_EINVAL, "avz_ibm_cov_keep_postmask_f32: n_fft 512, hop 128/256 only");
  return launch_cov512(mix, tgt, itf, nullptr, B, L, hop, 0.f, norm_eps, ibm_bits, R, msum, ws, spec,
                       (cudaStream_t)stream, 2);
}

int avz_wave_mask_cov_keep_f32(const float* mix, const float* mask, int B, int64_t L, int n_fft, int hop, float sqrt_eps,
                               float norm_eps, float* R, float* msum, void* ws, void* spec, void* stream) {
  if (!mix || !mask || !R || !msum || !ws || !spec || B <= 0 || B > 65535)
    return set_error(AVZ_EINVAL, "avz_wave_mask_cov_keep_f32: null pointer or empty batch");
  int rc = check_fft_args(n_fft, hop, L);
  if (rc) return rc;
  if (use_opt1024(n_fft, hop))
    return mask_cov1024(mix, mask, B, L, sqrt_eps, norm_eps, R, msum, ws, spec, (cudaStream_t)stream);
  if (!use_opt512(n_fft, hop))
    return set_error(AVZ_EINVAL, "avz_wave_mask_cov_keep_f32: n_fft 512 (hop 128/256) or 1024 (hop 512) only");
  return launch_cov512(mix, nullptr, nullptr, mask, B, L, hop, sqrt_eps, norm_eps, nullptr, R, msum, ws, spec,
                       (cudaStream_t)stream);
}

// Post-filter of pass B as a kernel gain mode, with the buffer the mode needs checked.
static int gain_mode_of(const AvzMvdrCfg* cfg, const uint32_t* ibm_bits, const float* mask, int* gain) {
  switch (cfg->post_mode) {
    case AVZ_POST_NONE: *gain = GAIN_NONE; return AVZ_OK;
    case AVZ_POST_ONE_MINUS_NOISE:
      if (!ibm_bits) return set_error(AVZ_EINVAL, "AVZ_POST_ONE_MINUS_NOISE needs ibm_bits");
      *gain = GAIN_BITS;
      return AVZ_OK;
    case AVZ_POST_FLOOR:
    case AVZ_POST_MASK:
      if (!mask) return set_error(AVZ_EINVAL, "AVZ_POST_FLOOR / AVZ_POST_MASK need a float mask");
      *gain = cfg->post_mode == AVZ_POST_FLOOR ? GAIN_FLOOR : GAIN_MASK;
      return AVZ_OK;
    default: return set_error(AVZ_EINVAL, "post_mode=%d unknown", cfg->post_mode);
  }
}

static int apply_kept(const void* spec, const float* w, const uint32_t* ibm_bits, const float* mask, int B, int64_t L,
                      int n_fft, int hop, const AvzMvdrCfg* cfg, float* out, float* peak, int fuse_norm, float peak_eps,
                      void* stream, int sparse = 0) {
  if (!spec || !w || !cfg || !out || B <= 0 || B > 65535)
    return set_error(AVZ_EINVAL, "avz_mvdr_apply_kept_f32: null pointer or empty batch");
  int rc = check_fft_args(n_fft, hop, L);
  if (rc) return rc;
  const bool fast1024 = use_opt1024(n_fft, hop);
  if (!fast1024 && !use_opt512(n_fft, hop))
    return set_error(AVZ_EINVAL, "avz_mvdr_apply_kept_f32: n_fft 512 (hop 128/256) or 1024 (hop 512) only");
  // A float-mask post-filter without a mask pointer means: the mask the learned-mask pass A (avz_wave_mask_cov_keep_f32)
  // re-laid behind the kept spectrum in this same `spec` - pass B then skips its own transposition (n_fft 512 only).
  const bool needs_mask = cfg->post_mode == AVZ_POST_FLOOR || cfg->post_mode == AVZ_POST_MASK;
  const int staged = (needs_mask && !mask && !fast1024) ? 1 : 0;
  int gain = GAIN_NONE;
  if (staged) {
    gain = cfg->post_mode == AVZ_POST_FLOOR ? GAIN_FLOOR : GAIN_MASK;
  } else {
    rc = gain_mode_of(cfg, ibm_bits, mask, &gain);
    if (rc) return rc;
  }
  if (fast1024) {
    if (gain == GAIN_BITS || fuse_norm)
      return set_error(AVZ_EINVAL, "avz_mvdr_apply_kept_f32: n_fft 1024 takes float masks only, no fused normalisation");
    return o1024::launch_apply(nullptr, spec, w, mask, gain, cfg->post_floor, B, L, out, peak, (cudaStream_t)stream);
  }
  if (sparse && (fast1024 || gain != GAIN_BITS))
    return set_error(AVZ_EINVAL, "a sparse kept spectrum goes with AVZ_POST_ONE_MINUS_NOISE and its ibm_bits (n_fft 512)");
  return (hop == 128) ? o512::launch_apply<128>(nullptr, spec, w, ibm_bits, mask, gain, cfg->post_floor, B, L, out, peak,
                                                fuse_norm, peak_eps, staged, (cudaStream_t)stream, sparse)
                      : o512::launch_apply<256>(nullptr, spec, w, ibm_bits, mask, gain, cfg->post_floor, B, L, out, peak,
                                                fuse_norm, peak_eps, staged, (cudaStream_t)stream, sparse);
}

int avz_mvdr_apply_kept_sparse_f32(const void* spec, const float* w, const uint32_t* ibm_bits, int B, int64_t L, int n_fft,
                                   int hop, const AvzMvdrCfg* cfg, float* out, float* peak, void* stream) {
  return apply_kept(spec, w, ibm_bits, nullptr, B, L, n_fft, hop, cfg, out, peak, 0, 0.f, stream, 1);
}

int avz_mvdr_apply_kept_f32(const void* spec, const float* w, const uint32_t* ibm_bits, const float* mask, int B,
                            int64_t L, int n_fft, int hop, const AvzMvdrCfg* cfg, float* out, float* peak, void* stream) {
  return apply_kept(spec, w, ibm_bits, mask, B, L, n_fft, hop, cfg, out, peak, 0, 0.f, stream);
}

int avz_mvdr_apply_kept_norm_f32(const void* spec, const float* w, const uint32_t* ibm_bits, const float* mask, int B,
                                 int64_t L, int n_fft, int hop, const AvzMvdrCfg* cfg, float peak_eps, float* out,
                                 float* peak, void* stream) {
  if (!peak) return set_error(AVZ_EINVAL, "avz_mvdr_apply_kept_norm_f32: peak buffer required");
  return apply_kept(spec, w, ibm_bits, mask, B, L, n_fft, hop, cfg, out, peak, 1, peak_eps, stream);
}

// ---- pass A with finalize + weights folded into its last block per utterance (n_fft 512 fast path)
int avz_ibm_cov_weights_keep_f32(const float* mix, const float* tgt, const float* itf, int B, int64_t L, int n_fft, int hop,
                                 const AvzMvdrCfg* cfg, const float* dvec, uint32_t* ibm_bits, float* R, float* msum,
                                 float* w, void* ws, void* spec, int sparse, void* stream) {
  if (!mix || !tgt || !itf || !cfg || !dvec || !ibm_bits || !R || !msum || !w || !ws || B <= 0 || B > 65535)
    return set_error(AVZ_EINVAL, "avz_ibm_cov_weights_keep_f32: null pointer or bad batch size");
  int rc = check_fft_args(n_fft, hop, L);
  if (rc) return rc;
  if (!use_opt512(n_fft, hop))
    return set_error(AVZ_EINVAL, "avz_ibm_cov_weights_keep_f32: n_fft 512 with hop 128 or 256 only");
  if (sparse < 0 || sparse > 2) return set_error(AVZ_EINVAL, "avz_ibm_cov_weights_keep_f32: sparse must be 0, 1 or 2");
  CovTailArgs tail{dvec, R, msum, w, cfg, cfg->norm_eps};
  int chunks = 0;
  return (hop == 128) ? o512::launch_ibm_cov<128>(mix, tgt, itf, nullptr, B, L, 0.f, ibm_bits, (float*)ws, &chunks, spec,
                                                  (cudaStream_t)stream, &tail, sparse)
                      : o512::launch_ibm_cov<256>(mix, tgt, itf, nullptr, B, L, 0.f, ibm_bits, (float*)ws, &chunks, spec,
                                                  (cudaStream_t)stream, &tail, sparse);
}

// ---- the whole oracle path in two small launches + one persistent kernel (n_fft 512, hop 128 / 256)
int64_t avz_oracle_fused_ws_bytes(int B, int64_t L, int n_fft, int hop) {
  if (B <= 0 || B > 65535 || check_fft_args(n_fft, hop, L) || !use_opt512(n_fft, hop)) return 0;
  return o512::fused_ws_bytes(B, L, hop);
}

int avz_oracle_fused_f32(const float* mix, const float* tgt, const float* itf, int B, int64_t L, int n_fft, int hop,
                         const AvzMvdrCfg* cfg, float peak_eps, const float* dvec, uint32_t* ibm_bits, float* R, float* msum,
                         float* w, float* out, float* peak, void* ws, void* stream) {
  if (!mix || !tgt || !itf || !cfg || !dvec || !ibm_bits || !R || !msum || !w || !out || !peak || !ws || B <= 0 || B > 65535)
    return set_error(AVZ_EINVAL, "avz_oracle_fused_f32: null pointer or bad batch size");
  int rc = check_fft_args(n_fft, hop, L);
  if (rc) return rc;
  if (!use_opt512(n_fft, hop)) return set_error(AVZ_EINVAL, "avz_oracle_fused_f32: n_fft 512 with hop 128 or 256 only");
  if (cfg->post_mode != AVZ_POST_ONE_MINUS_NOISE && cfg->post_mode != AVZ_POST_NONE)
    return set_error(AVZ_EINVAL, "avz_oracle_fused_f32: post-filter must be AVZ_POST_ONE_MINUS_NOISE or AVZ_POST_NONE");
  return (hop == 128) ? o512::launch_oracle_fused<128>(mix, tgt, itf, B, L, cfg, cfg->norm_eps, peak_eps, dvec, ibm_bits, R,
                                                       msum, w, out, peak, ws, (cudaStream_t)stream)
                      : o512::launch_oracle_fused<256>(mix, tgt, itf, B, L, cfg, cfg->norm_eps, peak_eps, dvec, ibm_bits, R,
                                                       msum, w, out, peak, ws, (cudaStream_t)stream);
}

// ---- streaming (n_fft 512 / hop 128): one hop per call for n_streams independent 2-mic streams ---------------
int64_t avz_stream_state_bytes(int n_streams) {
  return n_streams > 0 ? (int64_t)n_streams * (2 * 384 + 384 + 5 * 288) * (int64_t)sizeof(float) : 0;
}

int avz_stream_step_f32(float* state, const float* hop_in, const float* noise_w, const float* dvec, int n_streams, int t,
                        int t_end, float lambda, const AvzMvdrCfg* cfg, float* hop_out, void* stream) {
  if (!state || !hop_in || !dvec || !cfg || !hop_out || n_streams <= 0)
    return set_error(AVZ_EINVAL, "avz_stream_step_f32: null pointer or no streams");
  if (!(lambda >= 0.f && lambda < 1.f)) return set_error(AVZ_EINVAL, "avz_stream_step_f32: lambda must be in [0, 1)");
  return o512::launch_stream_step(state, hop_in, noise_w, dvec, n_streams, t, t_end, lambda, cfg, hop_out,
                                  (cudaStream_t)stream);
}

int avz_mvdr_apply_f32(const float* mix, const float* w, const uint32_t* ibm_bits, const float* mask, int B, int64_t L,
                       int n_fft, int hop, const AvzMvdrCfg* cfg, float* out, float* peak, void* stream) {
  if (!mix || !w || !cfg || !out || B <= 0 || B > 65535) return set_error(AVZ_EINVAL, "avz_mvdr_apply_f32: null pointer or empty batch");
  int rc = check_fft_args(n_fft, hop, L);
  if (rc) return rc;
  int gain = GAIN_NONE;
  rc = gain_mode_of(cfg, ibm_bits, mask, &gain);
  if (rc) return rc;
  const int T = (int)avz_num_frames(L, n_fft, hop);
  if (use_opt512(n_fft, hop)) {
    return (hop == 128) ? o512::launch_apply<128>(mix, nullptr, w, ibm_bits, mask, gain, cfg->post_floor, B, L, out, peak,
                                                  0, 0.f, 0, (cudaStream_t)stream)
                        : o512::launch_apply<256>(mix, nullptr, w, ibm_bits, mask, gain, cfg->post_floor, B, L, out, peak,
                                                  0, 0.f, 0, (cudaStream_t)stream);
  }
  if (use_opt1024(n_fft, hop) && gain != GAIN_BITS)
    return o1024::launch_apply(mix, nullptr, w, mask, gain, cfg->post_floor, B, L, out, peak, (cudaStream_t)stream);
  AVZ_DISPATCH_N(n_fft, (launch_synth<N_, SRC_MIX>(mix, nullptr, w, ibm_bits, mask, gain, cfg->post_floor, B, L, T, hop,
                                                   out, peak, (cudaStream_t)stream)));
}

}  // extern "C"

// ---- chunk drivers: 2 s windows at 50 % overlap read in place, count-averaged OLA ----
namespace avz {

// final[r][s] = sum over windows i covering s of outs[r * n_win + i][s - i * stride] / max(count, 1), windows taken in
// increasing i like the reference's `out_buf[start:start+w] += chunk_out[:w]` loop (full_audio.../inference.py:151-155,
// Final_pipeline/src/inference.py:225-233).  A window contributes its first `use_len` samples.  Deterministic gather.
__global__ void __launch_bounds__(256) k_chunk_ola(const float* __restrict__ outs, int n_win, int64_t olen, int64_t rec_len,
                                                   int stride, int64_t use_len, float* __restrict__ final_,
                                                   float* __restrict__ peak) {
  const int r = blockIdx.y;
  const float* o = outs + (int64_t)r * n_win * olen;
  float* f = final_ + (int64_t)r * rec_len;
  float m = 0.f;
  for (int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; s < rec_len; s += (int64_t)gridDim.x * blockDim.x) {
    int64_t i1 = s / stride;
    if (i1 > n_win - 1) i1 = n_win - 1;
    int64_t i0 = (s - use_len) / stride + 1;      // first i with i * stride + use_len > s
    if (s - use_len < 0) i0 = 0;
    float acc = 0.f, cnt = 0.f;
    for (int64_t i = i0; i <= i1; ++i) {
      acc += o[i * olen + (s - i * stride)];
      cnt += 1.f;
    }
    const float v = acc / (cnt > 0.f ? cnt : 1.f);
    f[s] = v;
    m = fmaxf(m, fabsf(v));
  }
  if (peak != nullptr) {
    m = warp_max(m);
    if ((threadIdx.x & 31) == 0) atomicMax(reinterpret_cast<unsigned int*>(peak + r), __float_as_uint(m));
  }
}

// WAV frames (frames x channels, interleaved PCM16: what soundfile.read hands the reference, oracle_debug.py:35-39) ->
// planar float32 [R][C][n] = pcm / 32768, the layout the fused kernels read.
__global__ void __launch_bounds__(256) k_pcm16_frames_to_planar(const int16_t* __restrict__ pcm, int64_t n, int C,
                                                                float* __restrict__ out) {
  const int r = blockIdx.y;
  const int16_t* p = pcm + (int64_t)r * n * C;
  float* o = out + (int64_t)r * C * n;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n * C; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t c = i / n, s = i - c * n;        // consecutive threads write consecutive samples of one channel
    o[i] = (float)p[s * C + c] * (1.0f / 32768.0f);
  }
}

static int check_chunk(const AvzChunkView* cv, int R, int64_t win, int n_fft, int hop, const char* who) {
  if (!cv || R <= 0 || cv->rec_len <= 0 || cv->stride <= 0 || cv->n_windows <= 0)
    return set_error(AVZ_EINVAL, "%s: bad chunk view", who);
  if ((int64_t)R * cv->n_windows > 65535) return set_error(AVZ_EINVAL, "%s: R * n_windows > 65535", who);
  if (cv->rec_len >= (1ll << 30)) return set_error(AVZ_EINVAL, "%s: recording too long", who);
  if (!use_opt1024(n_fft, hop))
    return set_error(AVZ_EINVAL, "%s: the in-place chunk view runs on the n_fft 1024 / hop 512 fast path (the STFT shape "
                                 "of every chunk driver in the reference)", who);
  return check_fft_args(n_fft, hop, win);
}

}  // namespace avz

extern "C" {

int avz_chunk_features_f32(const float* rec, int R, const AvzChunkView* cv, int64_t win, int n_fft, int hop, int mode,
                           float* X, void* stream) {
  if (!rec || !X || mode < 0 || mode > 2) return set_error(AVZ_EINVAL, "avz_chunk_features_f32: bad argument");
  int rc = check_chunk(cv, R, win, n_fft, hop, "avz_chunk_features_f32");
  if (rc) return rc;
  return o1024::launch_features(rec, R * cv->n_windows, win, mode, X, (cudaStream_t)stream, cv);
}

int avz_chunk_mask_cov_f32(const float* rec, const float* mask, int R, const AvzChunkView* cv, int64_t win, int n_fft,
                           int hop, float sqrt_eps, float norm_eps, float* Rcov, float* msum, void* ws, void* spec,
                           void* stream) {
  if (!rec || !mask || !Rcov || !msum || !ws) return set_error(AVZ_EINVAL, "avz_chunk_mask_cov_f32: null pointer");
  int rc = check_chunk(cv, R, win, n_fft, hop, "avz_chunk_mask_cov_f32");
  if (rc) return rc;
  return mask_cov1024(rec, mask, R * cv->n_windows, win, sqrt_eps, norm_eps, Rcov, msum, ws, spec, (cudaStream_t)stream, cv);
}

int avz_chunk_mvdr_apply_f32(const float* rec, const void* spec, const float* w, const float* mask, int R,
                             const AvzChunkView* cv, int64_t win, int n_fft, int hop, const AvzMvdrCfg* cfg, float* out,
                             float* peak, void* stream) {
  if ((!rec && !spec) || !w || !cfg || !out) return set_error(AVZ_EINVAL, "avz_chunk_mvdr_apply_f32: null pointer");
  int rc = check_chunk(cv, R, win, n_fft, hop, "avz_chunk_mvdr_apply_f32");
  if (rc) return rc;
  int gain = GAIN_NONE;
  rc = gain_mode_of(cfg, nullptr, mask, &gain);
  if (rc) return rc;
  return o1024::launch_apply(spec ? nullptr : rec, spec, w, mask, gain, cfg->post_floor, R * cv->n_windows, win, out, peak,
                             (cudaStream_t)stream, cv);
}

int avz_chunk_ola_f32(const float* outs, int R, int n_windows, int64_t olen, int64_t rec_len, int stride, int64_t use_len,
                      float* final_, float* peak, void* stream) {
  if (!outs || !final_ || R <= 0 || R > 65535 || n_windows <= 0 || olen <= 0 || rec_len <= 0 || stride <= 0 || use_len <= 0 ||
      use_len > olen)
    return set_error(AVZ_EINVAL, "avz_chunk_ola_f32: bad argument");
  int gx = (int)((rec_len + 1023) / 1024);
  if (gx > 148 * 8) gx = 148 * 8;
  k_chunk_ola<<<dim3(gx, R), 256, 0, (cudaStream_t)stream>>>(outs, n_windows, olen, rec_len, stride, use_len, final_, peak);
  AVZ_LAUNCH_OK("k_chunk_ola");
  return AVZ_OK;
}

int avz_pcm16_frames_to_planar_f32(const int16_t* pcm, int R, int64_t n, int C, float* out, void* stream) {
  if (!pcm || !out || R <= 0 || R > 65535 || n <= 0 || C <= 0) return set_error(AVZ_EINVAL, "avz_pcm16_frames_to_planar_f32: bad argument");
  int gx = (int)((n * C + 1023) / 1024);
  if (gx > 148 * 8) gx = 148 * 8;
  k_pcm16_frames_to_planar<<<dim3(gx, R), 256, 0, (cudaStream_t)stream>>>(pcm, n, C, out);
  AVZ_LAUNCH_OK("k_pcm16_frames_to_planar");
  return AVZ_OK;
}

}  // extern "C"

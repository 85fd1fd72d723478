// Float64 forms of the unfused operators.
//
// The reference computes in float64 / complex128 whenever its inputs are float64 (scipy.signal.stft follows the input
// dtype; R, w and S are allocated `dtype=complex`, oracle_debug.py:57,67).  Two of its call sites are ill-conditioned
// enough that a float32 STFT is visible in the output: masked_mvdr.main (sigma = 1e-7 on a near-rank-1 covariance,
// masked_mvdr.py:76-128) and hybrid_hard_null_bf (eigenvector + condition-number threshold + constraint solve,
// Final_pipeline/src/inference.py:28-98).  These kernels give those sites the reference's own precision: the Python
// wrappers in ops.py route float64 / complex128 inputs here, float32 / complex64 inputs to the float32 kernels, as
// numpy and scipy do.  None of this is on the throughput path (one warp per frame or per bin, shared-memory radix-2).
#include "avz_common.cuh"

namespace avz {
namespace f64 {

struct cdd {
  double x, y;
};
__device__ __forceinline__ cdd cmul(cdd a, cdd b) { return {a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x}; }
__device__ __forceinline__ cdd cdiv(cdd a, cdd b) {
  const double den = b.x * b.x + b.y * b.y;
  return {(a.x * b.x + a.y * b.y) / den, (a.y * b.x - a.x * b.y) / den};
}

constexpr int kWarps = 4;

// In-place radix-2 decimation-in-time transform of N = 2^logn points by one warp; `buf` holds the input in
// bit-reversed order and the result in natural order.  tw[k] = exp(-2 pi i k / N).
template <bool INV>
__device__ __forceinline__ void warp_fft_f64(double2* buf, int N, int logn, const double2* __restrict__ tw, int lane) {
  for (int s = 1; s <= logn; ++s) {
    const int half = 1 << (s - 1), step = N >> s;
    for (int j = lane; j < N / 2; j += kWarp) {
      const int k = j & (half - 1);
      const int i0 = ((j >> (s - 1)) << s) + k, i1 = i0 + half;
      double2 w = tw[k * step];
      if (INV) w.y = -w.y;
      const double2 a = buf[i0], b = buf[i1];
      const double2 t = make_double2(b.x * w.x - b.y * w.y, b.x * w.y + b.y * w.x);
      buf[i0] = make_double2(a.x + t.x, a.y + t.y);
      buf[i1] = make_double2(a.x - t.x, a.y - t.y);
    }
    __syncwarp();
  }
}

__device__ __forceinline__ int ilog2(int n) { return 31 - __clz(n); }
__device__ __forceinline__ int bitrev(int i, int logn) { return (int)(__brev((unsigned)i) >> (32 - logn)); }

// scipy.signal.stft (oracle_debug.py:42): frame t of x_ext = [0]*(N/2) ++ x ++ zeros, periodic Hann, / sum(w).
// One warp per (b, c, t).  x [B*C, L] -> Y [B*C, F, T] complex128.
// View: signal `sig` = (utterance b, channel c) with C channels; n_win > 1 reads window b % n_win of planar recording
// b / n_win in place (AvzChunkView), zero past rec_len.  Plain batch: n_win = 1, rec_len = L.
struct View {
  int C, n_win, stride;
  int64_t rec_len;
};
template <typename TIn>
__global__ void __launch_bounds__(kWarps * 32) k_stft_f64(const TIn* __restrict__ x, View vw, int64_t n_sig, int64_t L_full,
                                                          int T, int N, int hop, double inv_wsum, double2* __restrict__ Y,
                                                          Tables tb) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double2* buf = reinterpret_cast<double2*>(smem_raw) + (size_t)warp * N;
  const int logn = ilog2(N), F = N / 2 + 1;
  for (int64_t job = (int64_t)blockIdx.x * kWarps + warp; job < n_sig * T; job += (int64_t)gridDim.x * kWarps) {
    const int64_t sig = job / T;
    const int t = (int)(job - sig * T);
    const int64_t b = sig / vw.C, c = sig - b * vw.C;
    const int64_t r = b / vw.n_win, wi = b - r * vw.n_win;
    const int64_t w0 = wi * vw.stride;
    const TIn* xs = x + (r * vw.C + c) * vw.rec_len + w0;
    int64_t L = vw.rec_len - w0;
    L = L < 0 ? 0 : (L > L_full ? L_full : L);
    const int64_t start = (int64_t)t * hop - N / 2;
    for (int n = lane; n < N; n += kWarp) {
      const int64_t i = start + n;
      const double v = (i >= 0 && i < L) ? (double)xs[i] * tb.win_d[n] : 0.0;
      buf[bitrev(n, logn)] = make_double2(v, 0.0);
    }
    __syncwarp();
    warp_fft_f64<false>(buf, N, logn, tb.tw_d, lane);
    double2* yo = Y + sig * (int64_t)F * T + t;
    for (int k = lane; k < F; k += kWarp) {
      double2 v = buf[k];
      v.x *= inv_wsum;
      v.y *= inv_wsum;
      if (k == 0 || k == N / 2) v.y = 0.0;   // rfft of a real frame: exactly real at DC and Nyquist
      yo[(int64_t)k * T] = v;
    }
    __syncwarp();
  }
}

// scipy.signal.istft (oracle_debug.py:93), first half: xs = irfft(S[:, t]) * sum(w) * w  -> frames [B, T, N].
__global__ void __launch_bounds__(kWarps * 32) k_istft_frames_f64(const double2* __restrict__ S, int64_t B, int T, int N,
                                                                  double wsum_over_n, double* __restrict__ frames,
                                                                  Tables tb) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double2* buf = reinterpret_cast<double2*>(smem_raw) + (size_t)warp * N;
  const int logn = ilog2(N), F = N / 2 + 1;
  for (int64_t job = (int64_t)blockIdx.x * kWarps + warp; job < B * T; job += (int64_t)gridDim.x * kWarps) {
    const int64_t b = job / T;
    const int t = (int)(job - b * T);
    const double2* si = S + b * (int64_t)F * T + t;
    for (int k = lane; k < F; k += kWarp) {
      double2 v = si[(int64_t)k * T];
      if (k == 0 || k == N / 2) v.y = 0.0;   // pocketfft c2r ignores Im(DC) and Im(Nyquist)
      buf[bitrev(k, logn)] = v;
      if (k > 0 && k < N / 2) buf[bitrev(N - k, logn)] = make_double2(v.x, -v.y);
    }
    __syncwarp();
    warp_fft_f64<true>(buf, N, logn, tb.tw_d, lane);
    double* fo = frames + job * N;
    for (int n = lane; n < N; n += kWarp) fo[n] = buf[n].x * wsum_over_n * tb.win_d[n];
    __syncwarp();
  }
}

// second half: overlap-add in frame order, / sum of w^2 (guard 1e-10), trimmed by N/2 at both ends.
__global__ void k_istft_ola_f64(const double* __restrict__ frames, int64_t B, int T, int N, int hop,
                                double* __restrict__ x, Tables tb) {
  const int64_t n_out = (int64_t)(T - 1) * hop;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < B * n_out;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = idx / n_out, i = idx - b * n_out;
    const int64_t p = i + N / 2;
    int64_t t0 = (p - N + hop) / hop;   // ceil((p - N + 1) / hop) for p - N + 1 >= 0
    if (p - N + 1 <= 0) t0 = 0;
    int64_t t1 = p / hop;
    if (t1 > T - 1) t1 = T - 1;
    double acc = 0.0, nrm = 0.0;
    for (int64_t t = t0; t <= t1; ++t) {
      const int j = (int)(p - t * hop);
      const double w = tb.win_d[j];
      acc += frames[(b * T + t) * N + j];
      nrm += w * w;
    }
    x[idx] = acc / (nrm > 1e-10 ? nrm : 1.0);
  }
}

// x[b, :] /= max|x[b, :]| + eps  (masked_mvdr.py:128); one block per row.
__global__ void __launch_bounds__(256) k_peak_normalise_f64(double* __restrict__ x, int64_t n, double eps,
                                                            double* __restrict__ peak_out) {
  __shared__ double s_m[8];
  double* xb = x + (int64_t)blockIdx.x * n;
  double m = 0.0;
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) m = fmax(m, fabs(xb[i]));
  for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(kFull, m, o));
  if ((threadIdx.x & 31) == 0) s_m[threadIdx.x >> 5] = m;
  __syncthreads();
  m = 0.0;
  for (int i = 0; i < 8; ++i) m = fmax(m, s_m[i]);
  if (peak_out && threadIdx.x == 0) peak_out[blockIdx.x] = m;
  const double den = m + eps;
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) xb[i] = xb[i] / den;
}

// masked_mvdr.py:37-46
__global__ void k_geometric_mask_f64(const double2* __restrict__ Y, int64_t B, int64_t n_per_b, double* __restrict__ mask) {
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < B * n_per_b;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = idx / n_per_b, i = idx - b * n_per_b;
    const double2 u = Y[(2 * b) * n_per_b + i], v = Y[(2 * b + 1) * n_per_b + i];
    const double pa = atan2(u.y, u.x), pb = atan2(v.y, v.x);
    mask[idx] = (fabs(pa - pb) > 0.0) ? 1.0 : 0.01;
  }
}

// oracle_debug.py:56-64; one warp per (b, k).  R [B,F,4] = (R00, R11, Re R01, Im R01), msum [B,F].
// TW = double: nw is the noise weight itself; TW = float: nw is the target-probability mask, weight = 1 - mask
// (full_audio.../inference.py:102-103, evaluated in float64 like `1 - Mask` on a float32 array promoted by the product).
template <typename TW>
__global__ void k_spec_mask_cov_f64(const double2* __restrict__ Y, const TW* __restrict__ nw, int64_t BF, int F, int T,
                                    double sqrt_eps, double norm_eps, double* __restrict__ R, double* __restrict__ msum) {
  const int64_t bk = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (bk >= BF) return;
  const int lane = threadIdx.x & 31;
  const int64_t b = bk / F, k = bk - b * F;
  const double2* y0 = Y + ((b * 2 + 0) * F + k) * T;
  const double2* y1 = Y + ((b * 2 + 1) * F + k) * T;
  const TW* m = nw + bk * T;
  double s0 = 0, s1 = 0, sr = 0, si = 0, sm = 0;
  for (int t = lane; t < T; t += kWarp) {
    const double2 a = y0[t], c = y1[t];
    const double mm = sizeof(TW) == sizeof(float) ? (double)(1.0f - (float)m[t]) : (double)m[t], ms = mm + sqrt_eps;
    s0 += ms * (a.x * a.x + a.y * a.y);
    s1 += ms * (c.x * c.x + c.y * c.y);
    sr += ms * (a.x * c.x + a.y * c.y);
    si += ms * (a.y * c.x - a.x * c.y);
    sm += mm;
  }
  s0 = warp_sum(s0);
  s1 = warp_sum(s1);
  sr = warp_sum(sr);
  si = warp_sum(si);
  sm = warp_sum(sm);
  if (lane == 0) {
    const double den = sm + norm_eps;
    R[4 * bk + 0] = s0 / den;
    R[4 * bk + 1] = s1 / den;
    R[4 * bk + 2] = sr / den;
    R[4 * bk + 3] = si / den;
    msum[bk] = sm;
  }
}

// oracle_debug.py:68-79 (closed-form 2x2 solve); sigma, w_eps as doubles (the float fields of AvzMvdrCfg would round 1e-7).
__global__ void k_mvdr_weights_f64(const double* __restrict__ R, const double2* __restrict__ dvec, int64_t BF, int F,
                                   double sigma, double w_eps, int hp_bins, int hp_mode, double2* __restrict__ w) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= BF) return;
  const int k = (int)(idx % F);
  cdd w0 = {0, 0}, w1 = {0, 0};
  if (k < hp_bins && hp_mode != AVZ_HP_NONE) {
    if (hp_mode == AVZ_HP_MIC0) w0.x = 1.0;
  } else {
    const double a = R[4 * idx] + sigma, c = R[4 * idx + 1] + sigma;
    const cdd bb = {R[4 * idx + 2], R[4 * idx + 3]};
    const cdd d0 = {dvec[2 * k].x, dvec[2 * k].y}, d1 = {dvec[2 * k + 1].x, dvec[2 * k + 1].y};
    const double det = a * c - (bb.x * bb.x + bb.y * bb.y);
    if (det == 0.0 || !isfinite(det)) {
      w0.x = 1.0;
    } else {
      const cdd bd1 = cmul(bb, d1), cbd0 = cmul({bb.x, -bb.y}, d0);
      const cdd u0 = {(c * d0.x - bd1.x) / det, (c * d0.y - bd1.y) / det};
      const cdd u1 = {(a * d1.x - cbd0.x) / det, (a * d1.y - cbd0.y) / det};
      const cdd t0 = cmul({d0.x, -d0.y}, u0), t1 = cmul({d1.x, -d1.y}, u1);
      const cdd den = {t0.x + t1.x + w_eps, t0.y + t1.y};
      w0 = cdiv(u0, den);
      w1 = cdiv(u1, den);
    }
  }
  w[2 * idx] = make_double2(w0.x, w0.y);
  w[2 * idx + 1] = make_double2(w1.x, w1.y);
}

// Final_pipeline/src/inference.py:56-94, same closed form as k_hybrid_null_weights (avz_pointwise.cu) in and out of
// float64.  zero_cov_nan: a bin whose principal eigenvector has a zero first component (e.g. an all-zero covariance)
// yields NaN like the reference's v_int / (v_int[0] / (|v_int[0]| + 1e-10)); otherwise it keeps delay-and-sum.
template <typename TOut>   // double2, or float2: float64 arithmetic with the weights rounded once for the float32 pass B
__global__ void k_hybrid_null_weights_f64(const double* __restrict__ R, const double2* __restrict__ dvec, int64_t BF, int F,
                                          int bypass_bins, int zero_cov_nan, TOut* __restrict__ w) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= BF) return;
  const int k = (int)(idx % F);
  cdd w0 = {1.0, 0.0}, w1 = {0.0, 0.0};
  if (k >= bypass_bins) {
    const double a = R[4 * idx], c = R[4 * idx + 1];
    const cdd b = {R[4 * idx + 2], R[4 * idx + 3]};
    const cdd d0 = {dvec[2 * k].x, dvec[2 * k].y}, d1 = {dvec[2 * k + 1].x, dvec[2 * k + 1].y};
    const cdd den = {d0.x + 1e-10, d0.y};
    const cdd vt0 = cdiv(d0, den), vt1 = cdiv(d1, den);
    w0 = {0.5 * vt0.x, 0.5 * vt0.y};
    w1 = {0.5 * vt1.x, 0.5 * vt1.y};
    const double half = 0.5 * (a - c), bb = b.x * b.x + b.y * b.y;
    const double rad = sqrt(half * half + bb);
    cdd v0, v1;
    if (half > 0.0 || (half == 0.0 && bb > 0.0)) {
      v0 = {half + rad, 0.0};
      v1 = {b.x, -b.y};
    } else {   // a < c, or a == c with b == 0 (eigh returns the identity there: last column [0, 1])
      v0 = b;
      v1 = {rad - half, 0.0};
      if (half == 0.0 && bb == 0.0) v1 = {1.0, 0.0};
    }
    const double nrm = sqrt(v0.x * v0.x + v0.y * v0.y + v1.x * v1.x + v1.y * v1.y);
    const double m0 = sqrt(v0.x * v0.x + v0.y * v0.y) / (nrm > 0.0 ? nrm : 1.0);
    if (nrm > 0.0 && m0 > 0.0) {
      const cdd u0 = {v0.x / nrm, v0.y / nrm}, u1 = {v1.x / nrm, v1.y / nrm};
      const double s = m0 + 1e-10;
      const cdd vi0 = {s, 0.0};
      const cdd q = cdiv(u1, u0);
      const cdd vi1 = {q.x * s, q.y * s};
      const double g00 = vt0.x * vt0.x + vt0.y * vt0.y + vt1.x * vt1.x + vt1.y * vt1.y;
      const double g11 = vi0.x * vi0.x + vi1.x * vi1.x + vi1.y * vi1.y;
      const cdd g01 = {vt0.x * vi0.x + vt1.x * vi1.x + vt1.y * vi1.y, -vt0.y * vi0.x + vt1.x * vi1.y - vt1.y * vi1.x};
      const double Tr = g00 + g11, D = g00 * g11 - (g01.x * g01.x + g01.y * g01.y);
      const double disc = sqrt(fmax(Tr * Tr - 4.0 * D, 0.0));
      const double smax2 = 0.5 * (Tr + disc), smin2 = D / smax2;
      const bool ok = (D > 0.0) && (smax2 <= 100.0 * smin2);
      if (ok) {
        const cdd p = {vt0.x, -vt0.y}, qq = {vt1.x, -vt1.y}, rr = {vi0.x, -vi0.y}, ss = {vi1.x, -vi1.y};
        const cdd ps = cmul(p, ss), qr = cmul(qq, rr);
        const cdd det = {ps.x - qr.x, ps.y - qr.y};
        if (det.x != 0.0 || det.y != 0.0) {
          w0 = cdiv(ss, det);
          const cdd t = cdiv(rr, det);
          w1 = {-t.x, -t.y};
        }
      }
    } else if (zero_cov_nan) {
      const double qnan = nan("");
      w0 = {qnan, qnan};
      w1 = {qnan, qnan};
    }
  }
  w[2 * idx].x = w0.x;
  w[2 * idx].y = w0.y;
  w[2 * idx + 1].x = w1.x;
  w[2 * idx + 1].y = w1.y;
}

// oracle_debug.py:80
__global__ void k_beamform_f64(const double2* __restrict__ w, const double2* __restrict__ Y, int F, int T,
                               double2* __restrict__ S) {
  const int64_t bk = blockIdx.x;
  const int64_t b = bk / F, k = bk - b * F;
  const double2 w0 = w[2 * bk], w1 = w[2 * bk + 1];
  const double2* y0 = Y + ((b * 2 + 0) * F + k) * T;
  const double2* y1 = Y + ((b * 2 + 1) * F + k) * T;
  double2* s = S + bk * T;
  for (int t = blockIdx.y * blockDim.x + threadIdx.x; t < T; t += gridDim.y * blockDim.x) {
    const double2 a = y0[t], c = y1[t];
    // conj(w0) a + conj(w1) c
    s[t] = make_double2(w0.x * a.x + w0.y * a.y + w1.x * c.x + w1.y * c.y,
                        w0.x * a.y - w0.y * a.x + w1.x * c.y - w1.y * c.x);
  }
}

}  // namespace f64
}  // namespace avz

using namespace avz;

static int grid_for(int64_t n, int per_block, int cap) {
  int64_t g = (n + per_block - 1) / per_block;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

static double window_sum(int n_fft) {
  // scipy: win.sum() of the periodic Hann window evaluated in float64 (same table as the device's win_d)
  const double two_pi = 6.283185307179586476925286766559;
  double s = 0.0;
  for (int k = 0; k < n_fft; ++k) s += 0.5 - 0.5 * cos(two_pi * k / n_fft);
  return s;
}

extern "C" {

int avz_stft_f64(const double* x, int B, int C, int64_t L, int n_fft, int hop, double* Y, void* stream) {
  if (!x || !Y || B <= 0 || C <= 0) return set_error(AVZ_EINVAL, "avz_stft_f64: null pointer or empty batch");
  int rc = check_fft_args(n_fft, hop, L);
  if (rc) return rc;
  Tables tb;
  rc = tables_for(n_fft, &tb);
  if (rc) return rc;
  const int T = (int)avz_num_frames(L, n_fft, hop);
  const size_t smem = (size_t)f64::kWarps * n_fft * sizeof(double2);
  AVZ_CUDA_OK(cudaFuncSetAttribute(f64::k_stft_f64<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int64_t n_sig = (int64_t)B * C;
  f64::k_stft_f64<double><<<grid_for(n_sig * T, f64::kWarps, 148 * 64), f64::kWarps * 32, smem, (cudaStream_t)stream>>>(
      x, f64::View{C, 1, 0, L}, n_sig, L, T, n_fft, hop, 1.0 / window_sum(n_fft), reinterpret_cast<double2*>(Y), tb);
  AVZ_LAUNCH_OK("k_stft_f64");
  return AVZ_OK;
}

int64_t avz_wave_mask_cov_f64_ws_bytes(int B, int64_t L, int n_fft, int hop) {
  if (B <= 0 || check_fft_args(n_fft, hop, L)) return -1;
  return (int64_t)B * 2 * (n_fft / 2 + 1) * avz_num_frames(L, n_fft, hop) * (int64_t)sizeof(double2);
}

static int wave_mask_cov_f64(const float* mix, const float* mask, int B, const AvzChunkView* cv, int64_t L, int n_fft,
                             int hop, double sqrt_eps, double norm_eps, double* R, double* msum, void* ws, void* stream) {
  int rc = check_fft_args(n_fft, hop, L);
  if (rc) return rc;
  Tables tb;
  rc = tables_for(n_fft, &tb);
  if (rc) return rc;
  const int T = (int)avz_num_frames(L, n_fft, hop), F = n_fft / 2 + 1;
  const size_t smem = (size_t)f64::kWarps * n_fft * sizeof(double2);
  AVZ_CUDA_OK(cudaFuncSetAttribute(f64::k_stft_f64<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int64_t n_sig = (int64_t)B * 2;
  double2* Y = static_cast<double2*>(ws);
  const f64::View vw = cv ? f64::View{2, cv->n_windows, cv->stride, cv->rec_len} : f64::View{2, 1, 0, L};
  f64::k_stft_f64<float><<<grid_for(n_sig * T, f64::kWarps, 148 * 64), f64::kWarps * 32, smem, (cudaStream_t)stream>>>(
      mix, vw, n_sig, L, T, n_fft, hop, 1.0 / window_sum(n_fft), Y, tb);
  AVZ_LAUNCH_OK("k_stft_f64");
  const int wpb = 8;
  const int64_t BF = (int64_t)B * F;
  f64::k_spec_mask_cov_f64<float><<<(unsigned)((BF + wpb - 1) / wpb), wpb * 32, 0, (cudaStream_t)stream>>>(
      Y, mask, BF, F, T, sqrt_eps, norm_eps, R, msum);
  AVZ_LAUNCH_OK("k_spec_mask_cov_f64");
  return AVZ_OK;
}

int avz_wave_mask_cov_f64(const float* mix, const float* mask, int B, int64_t L, int n_fft, int hop, double sqrt_eps,
                          double norm_eps, double* R, double* msum, void* ws, void* stream) {
  if (!mix || !mask || !R || !msum || !ws || B <= 0) return set_error(AVZ_EINVAL, "avz_wave_mask_cov_f64: bad argument");
  return wave_mask_cov_f64(mix, mask, B, nullptr, L, n_fft, hop, sqrt_eps, norm_eps, R, msum, ws, stream);
}

int avz_chunk_mask_cov_f64(const float* rec, const float* mask, int R_, const AvzChunkView* cv, int64_t win, int n_fft,
                           int hop, double sqrt_eps, double norm_eps, double* Rcov, double* msum, void* ws, void* stream) {
  if (!rec || !mask || !Rcov || !msum || !ws || !cv || R_ <= 0 || cv->n_windows <= 0 || cv->stride <= 0 || cv->rec_len <= 0)
    return set_error(AVZ_EINVAL, "avz_chunk_mask_cov_f64: bad argument");
  return wave_mask_cov_f64(rec, mask, R_ * cv->n_windows, cv, win, n_fft, hop, sqrt_eps, norm_eps, Rcov, msum, ws, stream);
}

int64_t avz_istft_f64_ws_bytes(int B, int T, int n_fft) {
  if (B <= 0 || T < 2 || n_fft <= 0) return -1;
  return (int64_t)B * T * n_fft * (int64_t)sizeof(double);
}

int avz_istft_f64(const double* S, int B, int T, int n_fft, int hop, double* x, void* ws, void* stream) {
  if (!S || !x || !ws || B <= 0 || T < 2) return set_error(AVZ_EINVAL, "avz_istft_f64: null pointer, empty batch or T < 2");
  int rc = check_fft_args(n_fft, hop, n_fft);
  if (rc) return rc;
  Tables tb;
  rc = tables_for(n_fft, &tb);
  if (rc) return rc;
  const size_t smem = (size_t)f64::kWarps * n_fft * sizeof(double2);
  AVZ_CUDA_OK(cudaFuncSetAttribute(f64::k_istft_frames_f64, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  double* frames = static_cast<double*>(ws);
  f64::k_istft_frames_f64<<<grid_for((int64_t)B * T, f64::kWarps, 148 * 64), f64::kWarps * 32, smem, (cudaStream_t)stream>>>(
      reinterpret_cast<const double2*>(S), B, T, n_fft, window_sum(n_fft) / n_fft, frames, tb);
  AVZ_LAUNCH_OK("k_istft_frames_f64");
  const int64_t n = (int64_t)B * (T - 1) * hop;
  f64::k_istft_ola_f64<<<grid_for(n, 256, 148 * 32), 256, 0, (cudaStream_t)stream>>>(frames, B, T, n_fft, hop, x, tb);
  AVZ_LAUNCH_OK("k_istft_ola_f64");
  return AVZ_OK;
}

int avz_peak_normalise_f64(double* x, int B, int64_t n, double peak_eps, double* peak, void* stream) {
  if (!x || B <= 0 || n <= 0) return set_error(AVZ_EINVAL, "avz_peak_normalise_f64: bad argument");
  f64::k_peak_normalise_f64<<<B, 256, 0, (cudaStream_t)stream>>>(x, n, peak_eps, peak);
  AVZ_LAUNCH_OK("k_peak_normalise_f64");
  return AVZ_OK;
}

int avz_geometric_mask_f64(const double* Y, int B, int F, int T, double* mask, void* stream) {
  if (!Y || !mask || B <= 0 || F <= 0 || T <= 0) return set_error(AVZ_EINVAL, "avz_geometric_mask_f64: bad argument");
  const int64_t n = (int64_t)F * T;
  f64::k_geometric_mask_f64<<<grid_for((int64_t)B * n, 256, 148 * 16), 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const double2*>(Y), B, n, mask);
  AVZ_LAUNCH_OK("k_geometric_mask_f64");
  return AVZ_OK;
}

int avz_spec_mask_cov_f64(const double* Y, const double* noise_w, int B, int F, int T, double sqrt_eps, double norm_eps,
                          double* R, double* msum, void* stream) {
  if (!Y || !noise_w || !R || !msum || B <= 0 || F <= 0 || T <= 0)
    return set_error(AVZ_EINVAL, "avz_spec_mask_cov_f64: bad argument");
  const int wpb = 8;
  const int64_t BF = (int64_t)B * F;
  f64::k_spec_mask_cov_f64<double><<<(unsigned)((BF + wpb - 1) / wpb), wpb * 32, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const double2*>(Y), noise_w, BF, F, T, sqrt_eps, norm_eps, R, msum);
  AVZ_LAUNCH_OK("k_spec_mask_cov_f64");
  return AVZ_OK;
}

int avz_mvdr_weights_f64(const double* R, const double* dvec, int B, int F, double sigma, double w_eps, int hp_bins,
                         int hp_mode, double* w, void* stream) {
  if (!R || !dvec || !w || B <= 0 || F <= 0) return set_error(AVZ_EINVAL, "avz_mvdr_weights_f64: bad argument");
  const int64_t BF = (int64_t)B * F;
  f64::k_mvdr_weights_f64<<<(unsigned)((BF + 127) / 128), 128, 0, (cudaStream_t)stream>>>(
      R, reinterpret_cast<const double2*>(dvec), BF, F, sigma, w_eps, hp_bins, hp_mode, reinterpret_cast<double2*>(w));
  AVZ_LAUNCH_OK("k_mvdr_weights_f64");
  return AVZ_OK;
}

int avz_hybrid_null_weights_f64(const double* R, const double* dvec, int B, int F, int bypass_bins, int zero_cov_nan,
                                double* w, void* stream) {
  if (!R || !dvec || !w || B <= 0 || F <= 0 || bypass_bins < 0)
    return set_error(AVZ_EINVAL, "avz_hybrid_null_weights_f64: bad argument");
  const int64_t BF = (int64_t)B * F;
  f64::k_hybrid_null_weights_f64<double2><<<(unsigned)((BF + 127) / 128), 128, 0, (cudaStream_t)stream>>>(
      R, reinterpret_cast<const double2*>(dvec), BF, F, bypass_bins, zero_cov_nan, reinterpret_cast<double2*>(w));
  AVZ_LAUNCH_OK("k_hybrid_null_weights_f64");
  return AVZ_OK;
}

int avz_hybrid_null_weights_f64_w32(const double* R, const double* dvec, int B, int F, int bypass_bins, int zero_cov_nan,
                                    float* w, void* stream) {
  if (!R || !dvec || !w || B <= 0 || F <= 0 || bypass_bins < 0)
    return set_error(AVZ_EINVAL, "avz_hybrid_null_weights_f64_w32: bad argument");
  const int64_t BF = (int64_t)B * F;
  f64::k_hybrid_null_weights_f64<float2><<<(unsigned)((BF + 127) / 128), 128, 0, (cudaStream_t)stream>>>(
      R, reinterpret_cast<const double2*>(dvec), BF, F, bypass_bins, zero_cov_nan, reinterpret_cast<float2*>(w));
  AVZ_LAUNCH_OK("k_hybrid_null_weights_f64");
  return AVZ_OK;
}

int avz_beamform_f64(const double* w, const double* Y, int B, int F, int T, double* S, void* stream) {
  if (!w || !Y || !S || B <= 0 || F <= 0 || T <= 0) return set_error(AVZ_EINVAL, "avz_beamform_f64: bad argument");
  dim3 grid((unsigned)((int64_t)B * F), grid_for(T, 256, 8));
  f64::k_beamform_f64<<<grid, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const double2*>(w),
                                                              reinterpret_cast<const double2*>(Y), F, T,
                                                              reinterpret_cast<double2*>(S));
  AVZ_LAUNCH_OK("k_beamform_f64");
  return AVZ_OK;
}

}  // extern "C"

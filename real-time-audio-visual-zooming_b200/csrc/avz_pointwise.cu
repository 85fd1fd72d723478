// Per-bin / per-sample kernels of the mask-MVDR path: closed-form 2x2 MVDR weights, beamforming of a given
// spectrum, masks from given spectra, masked covariance of a given spectrum, features, projection scores.
// These are HBM-bound (or tiny); the arithmetic that decides parity (2x2 solve, score sums) is float64.
#include "avz_common.cuh"

namespace avz {

// ------------------------------------------------------------------------------------------
// MVDR weights.  Replaces oracle_debug.py:68-79 (np.linalg.solve on R + sigma I, then w / (d^H w + eps)).
// ------------------------------------------------------------------------------------------
__global__ void k_mvdr_weights(const float4* __restrict__ R, const float2* __restrict__ dvec, int B, int F,
                               AvzMvdrCfg cfg, float2* __restrict__ w) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * F) return;
  const int k = idx % F;
  float2 w0, w1;
  mvdr_weights_bin(R[idx], dvec[2 * k], dvec[2 * k + 1], k, cfg, w0, w1);
  w[2 * (int64_t)idx] = w0;
  w[2 * (int64_t)idx + 1] = w1;
}

// ------------------------------------------------------------------------------------------
// Hybrid hard-null weights.  Replaces Final_pipeline/src/inference.py:56-94 per bin, closed form in float64:
// principal eigenvector of the Hermitian 2x2 interference covariance (np.linalg.eigh), phase-normalised to mic 0;
// target steering vector normalised to mic 0; C = [v_tgt, v_int]; cond_2(C) > 10 -> delay-and-sum v_tgt / 2, else
// w solves C^H w = [1, 0].  Bins k < bypass_bins pass mic 0 (w = [1, 0]).
// ------------------------------------------------------------------------------------------
__global__ void k_hybrid_null_weights(const float4* __restrict__ R, const float2* __restrict__ dvec, int B, int F,
                                      int bypass_bins, int zero_cov_nan, float2* __restrict__ w) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * F) return;
  const int k = idx % F;
  cd w0 = {1.0, 0.0}, w1 = {0.0, 0.0};
  if (k >= bypass_bins) {
    const float4 r = R[idx];
    const double a = r.x, c = r.y;
    const cd b = {(double)r.z, (double)r.w};
    // target steering vector, normalised to mic 0: v / (v[0] + 1e-10)
    const cd d0 = {(double)dvec[2 * k].x, (double)dvec[2 * k].y};
    const cd d1 = {(double)dvec[2 * k + 1].x, (double)dvec[2 * k + 1].y};
    const cd den = {d0.x + 1e-10, d0.y};
    const cd vt0 = cddiv(d0, den), vt1 = cddiv(d1, den);
    w0 = {0.5 * vt0.x, 0.5 * vt0.y};   // delay-and-sum default
    w1 = {0.5 * vt1.x, 0.5 * vt1.y};
    // principal eigenvector of [[a, b], [conj b, c]]
    const double half = 0.5 * (a - c), bb = b.x * b.x + b.y * b.y;
    const double rad = sqrt(half * half + bb);
    cd v0, v1;
    if (half > 0.0 || (half == 0.0 && bb > 0.0)) {   // lambda - c = half + rad is the well-conditioned difference
      v0 = {half + rad, 0.0};
      v1 = {b.x, -b.y};
    } else {                    // lambda - a = rad - half; R00 == R11 with R01 == 0: eigh returns the identity, last column [0, 1]
      v0 = b;
      v1 = {(half == 0.0 && bb == 0.0) ? 1.0 : rad - half, 0.0};
    }
    const double nrm = sqrt(v0.x * v0.x + v0.y * v0.y + v1.x * v1.x + v1.y * v1.y);
    const double m0 = sqrt(v0.x * v0.x + v0.y * v0.y) / (nrm > 0.0 ? nrm : 1.0);   // |v_int[0]| of the unit vector
    if (nrm > 0.0 && m0 > 0.0) {
      // v_int = v / (v0 / (|v0| + 1e-10)):  v_int[0] = |v0| + 1e-10 (real), v_int[1] = v1 (|v0| + 1e-10) / v0
      const cd u0 = {v0.x / nrm, v0.y / nrm}, u1 = {v1.x / nrm, v1.y / nrm};
      const double s = m0 + 1e-10;
      const cd vi0 = {s, 0.0};
      const cd q = cddiv(u1, u0);
      const cd vi1 = {q.x * s, q.y * s};
      // cond_2 of C = [v_tgt, v_int] from the Gram matrix G = C^H C
      const double g00 = vt0.x * vt0.x + vt0.y * vt0.y + vt1.x * vt1.x + vt1.y * vt1.y;
      const double g11 = vi0.x * vi0.x + vi1.x * vi1.x + vi1.y * vi1.y;
      const cd g01 = {vt0.x * vi0.x + vt1.x * vi1.x + vt1.y * vi1.y, -vt0.y * vi0.x + vt1.x * vi1.y - vt1.y * vi1.x};
      const double T = g00 + g11, D = g00 * g11 - (g01.x * g01.x + g01.y * g01.y);
      const double disc = sqrt(fmax(T * T - 4.0 * D, 0.0));
      const double smax2 = 0.5 * (T + disc), smin2 = D / smax2;   // smin^2 = D / smax^2 avoids cancellation
      const bool ok = (D > 0.0) && (smax2 <= 100.0 * smin2);      // cond <= 10
      if (ok) {
        // C^H w = [1, 0]:  [[conj vt0, conj vt1], [conj vi0, conj vi1]] w = e1
        const cd p = {vt0.x, -vt0.y}, qq = {vt1.x, -vt1.y}, rr = {vi0.x, -vi0.y}, ss = {vi1.x, -vi1.y};
        const cd ps = cdmul(p, ss), qr = cdmul(qq, rr);
        const cd det = {ps.x - qr.x, ps.y - qr.y};
        if (det.x != 0.0 || det.y != 0.0) {
          w0 = cddiv(ss, det);
          const cd t = cddiv(rr, det);
          w1 = {-t.x, -t.y};
        }
      }
    }
    else if (zero_cov_nan) {
      // no principal direction with a non-zero mic-0 component (exactly zero covariance, or diagonal with R00 <= R11):
      // the reference divides by zero there and emits NaN; without the flag delay-and-sum is kept instead
      w0 = {nan(""), nan("")};
      w1 = {nan(""), nan("")};
    }
  }
  w[2 * (int64_t)idx] = make_float2((float)w0.x, (float)w0.y);
  w[2 * (int64_t)idx + 1] = make_float2((float)w1.x, (float)w1.y);
}

// S[b,k,t] = conj(w0) Y0 + conj(w1) Y1
__global__ void k_beamform(const float2* __restrict__ w, const float2* __restrict__ Y, int F, int T,
                           float2* __restrict__ S) {
  const int bk = blockIdx.x;  // b * F + k  (on grid.x: B * F exceeds the 65535 limit of grid.y at BASELINE batch sizes)
  const int b = bk / F, k = bk - b * F;
  const float2 w0 = w[2 * (int64_t)bk], w1 = w[2 * (int64_t)bk + 1];
  const float2* y0 = Y + (((int64_t)b * 2 + 0) * F + k) * T;
  const float2* y1 = Y + (((int64_t)b * 2 + 1) * F + k) * T;
  float2* s = S + (int64_t)bk * T;
  for (int t = blockIdx.y * blockDim.x + threadIdx.x; t < T; t += gridDim.y * blockDim.x) {
    const float2 a = y0[t], c = y1[t];
    // conj(w) * y = cmulc(y, w)
    s[t] = cadd(cmulc(a, w0), cmulc(c, w1));
  }
}

// out = (|a| > |b|) compared exactly: squares of float32 are exact in float64.
__global__ void k_mag_greater(const float2* __restrict__ a, const float2* __restrict__ b, int64_t n,
                              float* __restrict__ out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float2 u = a[i], v = b[i];
    const double pu = (double)u.x * u.x + (double)u.y * u.y;
    const double pv = (double)v.x * v.x + (double)v.y * v.y;
    out[i] = pu > pv ? 1.f : 0.f;
  }
}

// oracle_reverb.py:143-147: soft mask sqrt(Pt / (Pt + Pi + 1e-10)), P = |S|^2.
__global__ void k_irm(const float2* __restrict__ a, const float2* __restrict__ b, int64_t n, float* __restrict__ out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float pt = cabs2(a[i]), pi = cabs2(b[i]);
    out[i] = sqrtf(pt / (pt + pi + 1e-10f));
  }
}

// masked_mvdr.py:37-46: 1 where |angle(Y0) - angle(Y1)| > 0 else 0.01.  The difference of two atan2 values
// is zero exactly when the two angles are the same float64 number; float32 inputs are promoted first.
__global__ void k_geometric_mask(const float2* __restrict__ Y, int B, int64_t n_per_b, float* __restrict__ mask) {
  for (int b = blockIdx.y; b < B; b += gridDim.y) {   // grid.y is capped at 65535
    const float2* y0 = Y + (int64_t)b * 2 * n_per_b;
    const float2* y1 = y0 + n_per_b;
    float* m = mask + (int64_t)b * n_per_b;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_per_b; i += (int64_t)gridDim.x * blockDim.x) {
      const float2 u = y0[i], v = y1[i];
      const double pa = atan2((double)u.y, (double)u.x), pb = atan2((double)v.y, (double)v.x);
      m[i] = (fabs(pa - pb) > 0.0) ? 1.0f : 0.01f;
    }
  }
}

__global__ void k_ibm_unpack(const uint32_t* __restrict__ bits, int B, int F, int T, int FW, float* __restrict__ mask) {
  const int k = blockIdx.y;
  for (int b = blockIdx.z; b < B; b += gridDim.z) {   // grid.z is capped at 65535
    const uint32_t* bb = bits + (int64_t)b * T * FW + (k >> 5);
    float* m = mask + ((int64_t)b * F + k) * T;
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < T; t += gridDim.x * blockDim.x)
      m[t] = ((bb[(int64_t)t * FW] >> (k & 31)) & 1u) ? 1.f : 0.f;
  }
}

// R[b,k] = sum_t (m + sqrt_eps) y y^H / (sum_t m + norm_eps); one warp per (b, k), float64 accumulation.
__global__ void k_spec_mask_cov(const float2* __restrict__ Y, const float* __restrict__ nw, int B, int F, int T,
                                float sqrt_eps, float norm_eps, float4* __restrict__ R, float* __restrict__ msum) {
  const int warps_per_block = blockDim.x >> 5;
  const int bk = blockIdx.x * warps_per_block + (threadIdx.x >> 5);
  if (bk >= B * F) return;
  const int lane = threadIdx.x & 31;
  const int b = bk / F, k = bk - b * F;
  const float2* y0 = Y + (((int64_t)b * 2 + 0) * F + k) * T;
  const float2* y1 = Y + (((int64_t)b * 2 + 1) * F + k) * T;
  const float* m = nw + (int64_t)bk * T;
  double s0 = 0, s1 = 0, sr = 0, si = 0, sm = 0;
  for (int t = lane; t < T; t += kWarp) {
    const float2 a = y0[t], c = y1[t];
    const double mm = (double)m[t];
    const double ms = mm + (double)sqrt_eps;
    s0 += ms * ((double)a.x * a.x + (double)a.y * a.y);
    s1 += ms * ((double)c.x * c.x + (double)c.y * c.y);
    sr += ms * ((double)a.x * c.x + (double)a.y * c.y);
    si += ms * ((double)a.y * c.x - (double)a.x * c.y);
    sm += mm;
  }
  s0 = warp_sum(s0);
  s1 = warp_sum(s1);
  sr = warp_sum(sr);
  si = warp_sum(si);
  sm = warp_sum(sm);
  if (lane == 0) {
    const double inv = 1.0 / (sm + (double)norm_eps);
    R[bk] = make_float4((float)(s0 * inv), (float)(s1 * inv), (float)(sr * inv), (float)(si * inv));
    msum[bk] = (float)sm;
  }
}

// ------------------------------------------------------------------------------------------
// features from a spectrum (full_audio.../inference.py:91-94; Final_pipeline/src/inference.py:117-128)
// ------------------------------------------------------------------------------------------
__global__ void k_features(const float2* __restrict__ Y, int B, int F, int T, int mode, float* __restrict__ X) {
  const int k = blockIdx.y;
  for (int b = blockIdx.z; b < B; b += gridDim.z) {   // grid.z is capped at 65535
    const float2* y0 = Y + (((int64_t)b * 2 + 0) * F + k) * T;
    const float2* y1 = Y + (((int64_t)b * 2 + 1) * F + k) * T;
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < T; t += gridDim.x * blockDim.x) {
      float lm, ipd;
      feature_values(y0[t], y1[t], lm, ipd);
      store_features(X, mode, b, k, t, F, T, lm, ipd);
    }
  }
}

// ------------------------------------------------------------------------------------------
// projection scores (Final_pipeline/src/metrics.py:102-123, scripts/run_metrics.py:6-36), float64.
// One block per utterance; two passes so the residual is summed explicitly like the reference does.
// ------------------------------------------------------------------------------------------
template <int NV>
__device__ __forceinline__ void block_sum(double (&v)[NV], double* s_red /*[NV*32]*/) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
#pragma unroll
  for (int j = 0; j < NV; ++j) v[j] = warp_sum(v[j]);
  __syncthreads();
  if (lane == 0)
#pragma unroll
    for (int j = 0; j < NV; ++j) s_red[j * 32 + warp] = v[j];
  __syncthreads();
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    double s = 0;
    for (int i = 0; i < nw; ++i) s += s_red[j * 32 + i];
    v[j] = s;
  }
}

__global__ void __launch_bounds__(512) k_sir(const float* __restrict__ est, const float* __restrict__ tgt,
                                             const float* __restrict__ itf, int64_t n_est, int64_t n_ref,
                                             float4* __restrict__ scores) {
  __shared__ double s_red[5 * 32];
  const int b = blockIdx.x;
  const int64_t n = n_est < n_ref ? n_est : n_ref;
  const float* o = est + (int64_t)b * n_est;
  const float* t = tgt + (int64_t)b * n_ref;
  const float* i_ = itf + (int64_t)b * n_ref;
  double v[5] = {0, 0, 0, 0, 0};  // oo, tt, ii, ot, oi
  for (int64_t j = threadIdx.x; j < n; j += blockDim.x) {
    const double a = o[j], c = t[j], d = i_[j];
    v[0] += a * a;
    v[1] += c * c;
    v[2] += d * d;
    v[3] += a * c;
    v[4] += a * d;
  }
  block_sum<5>(v, s_red);
  const double eps = 1e-10;
  const double no = sqrt(v[0]), nt = sqrt(v[1]), ni = sqrt(v[2]);
  const double st = 1.0 / (nt + eps), si = 1.0 / (ni + eps);  // target / interferer scaled to unit norm
  const double alpha = v[3] * st, beta = v[4] * si;           // <o, t^>, <o, i^>
  double r[3] = {0, 0, 0};                                    // |e_t|^2, |e_i|^2, |e_n|^2
  for (int64_t j = threadIdx.x; j < n; j += blockDim.x) {
    const double et = alpha * ((double)t[j] * st), ei = beta * ((double)i_[j] * si);
    const double en = (double)o[j] - et - ei;
    r[0] += et * et;
    r[1] += ei * ei;
    r[2] += en * en;
  }
  block_sum<3>(r, s_red);
  if (threadIdx.x == 0) {
    const double osinr = 10.0 * log10(r[0] / (r[1] + r[2] + eps));
    const double osir = 10.0 * log10(r[0] / (r[1] + eps));
    // run_metrics.py also scales the output to unit norm: every power shrinks by so^2
    const double so = 1.0 / (no + eps), so2 = so * so;
    const double pt = r[0] * so2, pi = r[1] * so2 + 1e-10, pn = r[2] * so2 + 1e-10;
    const double sdr = 10.0 * log10(pt / (pi + pn));
    const double sir = 10.0 * log10(pt / pi);
    scores[b] = make_float4((float)osinr, (float)osir, (float)sdr, (float)sir);
  }
}

}  // namespace avz

using namespace avz;

static inline int grid1d(int64_t n, int block, int cap) {
  int64_t g = (n + block - 1) / block;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

extern "C" {

int avz_mvdr_weights_f32(const float* R, const float* dvec, int B, int F, const AvzMvdrCfg* cfg, float* w, void* stream) {
  if (!R || !dvec || !cfg || !w || B <= 0 || F <= 0) return set_error(AVZ_EINVAL, "avz_mvdr_weights_f32: bad argument");
  prof_begin(PROF_WEIGHTS, (cudaStream_t)stream);
  k_mvdr_weights<<<(B * F + 127) / 128, 128, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const float4*>(R), reinterpret_cast<const float2*>(dvec), B, F, *cfg, reinterpret_cast<float2*>(w));
  prof_end(PROF_WEIGHTS, (cudaStream_t)stream);
  AVZ_LAUNCH_OK("k_mvdr_weights");
  return AVZ_OK;
}

int avz_hybrid_null_weights_f32(const float* R, const float* dvec, int B, int F, int bypass_bins, int zero_cov_nan, float* w,
                                void* stream) {
  if (!R || !dvec || !w || B <= 0 || F <= 0 || bypass_bins < 0)
    return set_error(AVZ_EINVAL, "avz_hybrid_null_weights_f32: bad argument");
  k_hybrid_null_weights<<<(B * F + 127) / 128, 128, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const float4*>(R), reinterpret_cast<const float2*>(dvec), B, F, bypass_bins, zero_cov_nan,
      reinterpret_cast<float2*>(w));
  AVZ_LAUNCH_OK("k_hybrid_null_weights");
  return AVZ_OK;
}

int avz_beamform_f32(const float* w, const float* Y, int B, int F, int T, float* S, void* stream) {
  if (!w || !Y || !S || B <= 0 || F <= 0 || T <= 0) return set_error(AVZ_EINVAL, "avz_beamform_f32: bad argument");
  dim3 grid(B * F, grid1d(T, 256, 8));
  k_beamform<<<grid, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const float2*>(w), reinterpret_cast<const float2*>(Y),
                                                     F, T, reinterpret_cast<float2*>(S));
  AVZ_LAUNCH_OK("k_beamform");
  return AVZ_OK;
}

int avz_mag_greater_f32(const float* a, const float* b, int64_t n, float* out, void* stream) {
  if (!a || !b || !out || n <= 0) return set_error(AVZ_EINVAL, "avz_mag_greater_f32: bad argument");
  k_mag_greater<<<grid1d(n, 256, 148 * 16), 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const float2*>(a), reinterpret_cast<const float2*>(b), n, out);
  AVZ_LAUNCH_OK("k_mag_greater");
  return AVZ_OK;
}

int avz_irm_f32(const float* s_tgt, const float* s_int, int64_t n, float* out, void* stream) {
  if (!s_tgt || !s_int || !out || n <= 0) return set_error(AVZ_EINVAL, "avz_irm_f32: bad argument");
  k_irm<<<grid1d(n, 256, 148 * 16), 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const float2*>(s_tgt), reinterpret_cast<const float2*>(s_int), n, out);
  AVZ_LAUNCH_OK("k_irm");
  return AVZ_OK;
}

int avz_geometric_mask_f32(const float* Y, int B, int F, int T, float* mask, void* stream) {
  if (!Y || !mask || B <= 0 || F <= 0 || T <= 0) return set_error(AVZ_EINVAL, "avz_geometric_mask_f32: bad argument");
  const int64_t n = (int64_t)F * T;
  dim3 grid(grid1d(n, 256, 148 * 4), B < 65535 ? B : 65535);
  k_geometric_mask<<<grid, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const float2*>(Y), B, n, mask);
  AVZ_LAUNCH_OK("k_geometric_mask");
  return AVZ_OK;
}

int avz_ibm_unpack_f32(const uint32_t* ibm_bits, int B, int F, int T, float* mask, void* stream) {
  if (!ibm_bits || !mask || B <= 0 || F <= 0 || T <= 0) return set_error(AVZ_EINVAL, "avz_ibm_unpack_f32: bad argument");
  dim3 grid(grid1d(T, 128, 8), F, B < 65535 ? B : 65535);
  k_ibm_unpack<<<grid, 128, 0, (cudaStream_t)stream>>>(ibm_bits, B, F, T, (F + 31) / 32, mask);
  AVZ_LAUNCH_OK("k_ibm_unpack");
  return AVZ_OK;
}

int avz_spec_mask_cov_f32(const float* Y, const float* noise_w, int B, int F, int T, float sqrt_eps, float norm_eps,
                          float* R, float* msum, void* stream) {
  if (!Y || !noise_w || !R || !msum || B <= 0 || F <= 0 || T <= 0)
    return set_error(AVZ_EINVAL, "avz_spec_mask_cov_f32: bad argument");
  const int wpb = 8;
  k_spec_mask_cov<<<(B * F + wpb - 1) / wpb, wpb * 32, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const float2*>(Y), noise_w, B, F, T, sqrt_eps, norm_eps, reinterpret_cast<float4*>(R), msum);
  AVZ_LAUNCH_OK("k_spec_mask_cov");
  return AVZ_OK;
}

int avz_features_f32(const float* Y, int B, int F, int T, int mode, float* X, void* stream) {
  if (!Y || !X || B <= 0 || F <= 1 || T <= 0 || mode < 0 || mode > 2)
    return set_error(AVZ_EINVAL, "avz_features_f32: bad argument");
  dim3 grid(grid1d(T, 128, 8), F, B < 65535 ? B : 65535);
  k_features<<<grid, 128, 0, (cudaStream_t)stream>>>(reinterpret_cast<const float2*>(Y), B, F, T, mode, X);
  AVZ_LAUNCH_OK("k_features");
  return AVZ_OK;
}

int avz_sir_f32(const float* est, const float* tgt, const float* itf, int B, int64_t n_est, int64_t n_ref, float* scores,
                void* stream) {
  if (!est || !tgt || !itf || !scores || B <= 0 || n_est <= 0 || n_ref <= 0)
    return set_error(AVZ_EINVAL, "avz_sir_f32: bad argument");
  k_sir<<<B, 512, 0, (cudaStream_t)stream>>>(est, tgt, itf, n_est, n_ref, reinterpret_cast<float4*>(scores));
  AVZ_LAUNCH_OK("k_sir");
  return AVZ_OK;
}

}  // extern "C"

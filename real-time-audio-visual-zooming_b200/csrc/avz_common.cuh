// Shared device/host helpers for libavzoom (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cstdio>
#include <stdint.h>
#include <math.h>

#include "avzoom.h"

namespace avz {

constexpr int kWarp = 32;
constexpr unsigned kFull = 0xffffffffu;

// Warp index of this thread within its CTA, in a form the compiler's uniformity analysis accepts as warp-uniform
// (a broadcast from lane 0).  `threadIdx.x >> 5` is just as uniform but the compiler does not know it: every branch and
// loop bound derived from it counts as divergent, and each __shfl_sync / __ballot_sync inside such a region is then
// wrapped in WARPSYNC.COLLECTIVE ... ENDCOLLECTIVE with its operands moved through fixed registers (5 instructions
// and a serialising register dependency per shuffle instead of 1 instruction).
__device__ __forceinline__ int warp_id_uniform() { return __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0); }

// Per-(device, n_fft) constant tables, created once by tables_for().
struct Tables {
  const float2* tw;     // [n]  exp(-2 pi i k / n), rounded from float64
  const float* win;     // [n]  periodic Hann, rounded from float64
  const double2* tw_d;  // [n]  same twiddles in float64 (exact-tie path of the IBM)
  const double* win_d;  // [n]
};

// what pass A needs to also finalize the covariance and solve for the weights (k512_cov_w)
struct CovTailArgs {
  const float* dvec;
  float* R;
  float* msum;
  float* w;
  const AvzMvdrCfg* cfg;
  float norm_eps;
};

int set_error(int code, const char* fmt, ...);
int tables_for(int n_fft, Tables* out);
int check_fft_args(int n_fft, int hop, int64_t L);
int num_sms();

// Optional per-kernel timing for bench.py (avz_profile_enable / avz_profile_get): CUDA events recorded on the
// launching stream right before and after each kernel of the fused path.
enum ProfSlot { PROF_IBM = 0, PROF_FIXUP, PROF_COV, PROF_FINALIZE, PROF_WEIGHTS, PROF_APPLY, PROF_NORMALISE, PROF_COUNT };
void prof_begin(int slot, cudaStream_t st);
void prof_end(int slot, cudaStream_t st);

#define AVZ_CUDA_OK(expr)                                                                   \
  do {                                                                                      \
    cudaError_t _e = (expr);                                                                \
    if (_e != cudaSuccess) return avz::set_error(AVZ_ECUDA, "%s: %s", #expr, cudaGetErrorString(_e)); \
  } while (0)

#define AVZ_LAUNCH_OK(name)                                                                 \
  do {                                                                                      \
    cudaError_t _e = cudaGetLastError();                                                    \
    if (_e != cudaSuccess) return avz::set_error(AVZ_ECUDA, "launch %s: %s", name, cudaGetErrorString(_e)); \
  } while (0)

// ------------------------------------------------------------------------------------------
// complex helpers
// ------------------------------------------------------------------------------------------
// Packed float32 pairs (Blackwell sm_100: add/mul/fma.rn.f32x2 -> SASS FADD2 / FMUL2 / FFMA2).  A complex value
// (re, im) is one 64-bit register pair; one packed instruction does the work of two scalar ones in ONE issue slot
// (the FMA pipe still spends two cycles on it - measured, profiles/r2_f32x2_probe.md - so it is issue slots, not
// flops, that are saved: exactly what an issue-bound FFT needs).  ptxas folds the pack / unpack moves below into
// operand modifiers of the packed instruction: half swap (.LO_HI), per-half negation (.NP / .PN) and 32-bit
// broadcast (Rn.F32), so multiplying by +-i, conjugating and scaling by a real cost no instruction of their own.
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk2(float a, float b) { u64 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ u64 pk2(float2 a) { return pk2(a.x, a.y); }
__device__ __forceinline__ float2 up2(u64 v) { float2 r; asm("mov.b64 {%0,%1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v)); return r; }
__device__ __forceinline__ u64 add2(u64 a, u64 b) { u64 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ u64 mul2(u64 a, u64 b) { u64 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }

__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return up2(add2(pk2(a), pk2(b))); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return up2(add2(pk2(a), pk2(-b.x, -b.y))); }
// a + conj(b), a - conj(b)
__device__ __forceinline__ float2 caddc(float2 a, float2 b) { return up2(add2(pk2(a), pk2(b.x, -b.y))); }
__device__ __forceinline__ float2 csubc(float2 a, float2 b) { return up2(add2(pk2(a), pk2(-b.x, b.y))); }
// a * s, a * s + c for a real s
__device__ __forceinline__ float2 cscale(float2 a, float s) { return up2(mul2(pk2(a), pk2(s, s))); }
__device__ __forceinline__ float2 cfma_real(float2 a, float s, float2 c) { return up2(fma2(pk2(a), pk2(s, s), pk2(c))); }
// a * b = (a.x, a.y) b.x + (-a.y, a.x) b.y
__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
  return up2(fma2(pk2(-a.y, a.x), pk2(b.y, b.y), mul2(pk2(a), pk2(b.x, b.x))));
}
// a * conj(b) = (a.x, a.y) b.x + (a.y, -a.x) b.y
__device__ __forceinline__ float2 cmulc(float2 a, float2 b) {
  return up2(fma2(pk2(a.y, -a.x), pk2(b.y, b.y), mul2(pk2(a), pk2(b.x, b.x))));
}
__device__ __forceinline__ float cabs2(float2 a) { return fmaf(a.x, a.x, a.y * a.y); }

template <typename T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(kFull, v, o));
  return v;
}

// ------------------------------------------------------------------------------------------
// MVDR weights of one bin.  Replaces oracle_debug.py:68-79 (np.linalg.solve on R + sigma I, then w / (d^H w + eps)):
// closed-form 2x2 solve in float64; an exactly singular (or non-finite) system gives w = [1, 0] like the reference's
// LinAlgError branch; bins below the high-pass get 0 (AVZ_HP_ZERO) or [1, 0] (AVZ_HP_MIC0).
// ------------------------------------------------------------------------------------------
struct cd {
  double x, y;
};
__device__ __forceinline__ cd cdmul(cd a, cd b) { return {a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x}; }
__device__ __forceinline__ cd cddiv(cd a, cd b) {
  const double den = b.x * b.x + b.y * b.y;
  return {(a.x * b.x + a.y * b.y) / den, (a.y * b.x - a.x * b.y) / den};
}
__device__ __forceinline__ void mvdr_weights_bin(float4 r, float2 dv0, float2 dv1, int k, const AvzMvdrCfg& cfg, float2& w0,
                                                 float2& w1) {
  w0 = make_float2(0.f, 0.f);
  w1 = make_float2(0.f, 0.f);
  if (k < cfg.hp_bins && cfg.hp_mode != AVZ_HP_NONE) {
    if (cfg.hp_mode == AVZ_HP_MIC0) w0.x = 1.f;  // pass mic 0 through
    return;
  }
  const double a = (double)r.x + (double)cfg.sigma, c = (double)r.y + (double)cfg.sigma;
  const cd bb = {(double)r.z, (double)r.w};  // R01; R10 = conj
  const cd d0 = {(double)dv0.x, (double)dv0.y};
  const cd d1 = {(double)dv1.x, (double)dv1.y};
  const double det = a * c - (bb.x * bb.x + bb.y * bb.y);
  if (det == 0.0 || !isfinite(det)) {
    w0.x = 1.f;  // LinAlgError fallback w = [1, 0] (oracle_debug.py:78-79)
    return;
  }
  // u = inv([[a, b],[conj b, c]]) d
  const cd bd1 = cdmul(bb, d1);
  const cd cbd0 = cdmul({bb.x, -bb.y}, d0);
  const cd u0 = {(c * d0.x - bd1.x) / det, (c * d0.y - bd1.y) / det};
  const cd u1 = {(a * d1.x - cbd0.x) / det, (a * d1.y - cbd0.y) / det};
  // denom = d^H u + w_eps
  const cd t0 = cdmul({d0.x, -d0.y}, u0);
  const cd t1 = cdmul({d1.x, -d1.y}, u1);
  const cd den = {t0.x + t1.x + (double)cfg.w_eps, t0.y + t1.y};
  const cd q0 = cddiv(u0, den), q1 = cddiv(u1, den);
  w0 = make_float2((float)q0.x, (float)q0.y);
  w1 = make_float2((float)q1.x, (float)q1.y);
}

// ------------------------------------------------------------------------------------------
// Generic warp-level FFT in shared memory (any power-of-two N >= 64).  One warp transforms one
// length-N complex sequence with Stockham autosort passes (radix 4, plus one radix-2 pass when
// log2 N is odd), ping-ponging between two N-element buffers.  Returns the buffer that holds the
// result in natural order.  `tw` is exp(-2 pi i k / N) (shared or global memory).
// The specialised register-resident 512-point transform lives in avz_fft512.cuh.
// ------------------------------------------------------------------------------------------
template <int N, bool INV>
__device__ __forceinline__ float2* warp_fft_smem(float2* a, float2* b, const float2* __restrict__ tw, int lane) {
  float2* in = a;
  float2* out = b;
  int Ns = 1;
#pragma unroll 1
  for (; Ns * 4 <= N; Ns *= 4) {
    const int tstep = N / (4 * Ns);
    for (int j = lane; j < N / 4; j += kWarp) {
      const int k = j & (Ns - 1);
      float2 v0 = in[j];
      float2 v1 = in[j + N / 4];
      float2 v2 = in[j + N / 2];
      float2 v3 = in[j + 3 * N / 4];
      if (Ns > 1) {
        float2 w1 = tw[k * tstep];
        float2 w2 = tw[2 * k * tstep];
        float2 w3 = tw[3 * k * tstep];
        if (INV) { w1.y = -w1.y; w2.y = -w2.y; w3.y = -w3.y; }
        v1 = cmul(v1, w1);
        v2 = cmul(v2, w2);
        v3 = cmul(v3, w3);
      }
      const float2 t0 = cadd(v0, v2);
      const float2 t1 = csub(v0, v2);
      const float2 t2 = cadd(v1, v3);
      float2 t3 = csub(v1, v3);
      t3 = INV ? make_float2(-t3.y, t3.x) : make_float2(t3.y, -t3.x);  // * (+i) inverse, * (-i) forward
      const int base = (j - k) * 4 + k;
      out[base] = cadd(t0, t2);
      out[base + Ns] = cadd(t1, t3);
      out[base + 2 * Ns] = csub(t0, t2);
      out[base + 3 * Ns] = csub(t1, t3);
    }
    __syncwarp();
    float2* tmp = in; in = out; out = tmp;
  }
  if (Ns < N) {  // one radix-2 pass, Ns == N/2
    for (int j = lane; j < N / 2; j += kWarp) {
      float2 w1 = tw[j];  // k = j (Ns = N/2), angle -2 pi k / N
      if (INV) w1.y = -w1.y;
      const float2 v0 = in[j];
      const float2 v1 = cmul(in[j + N / 2], w1);
      out[j] = cadd(v0, v1);
      out[j + N / 2] = csub(v0, v1);
    }
    __syncwarp();
    float2* tmp = in; in = out; out = tmp;
  }
  return in;
}

// Load frame t of a pair of real signals (sb may be null) as one complex sequence, windowed:
// buf[n] = win[n] * scale * (sa[i] + i sb[i]),  i = t*hop - N/2 + n, zero outside [0, L).
// This is scipy's zero extension (boundary='zeros') and tail padding (padded=True).
// Returns this lane's share of sum |buf|^2 (for the IBM tie tolerance).
template <int N>
__device__ __forceinline__ float load_frame_pair(float2* buf, const float* __restrict__ sa, const float* __restrict__ sb,
                                                 int64_t L, int64_t start, const float* __restrict__ win, float scale,
                                                 int lane) {
  float e = 0.f;
#pragma unroll 4
  for (int n = lane; n < N; n += kWarp) {
    const int64_t i = start + n;
    float xa = 0.f, xb = 0.f;
    if (i >= 0 && i < L) {
      xa = __ldg(sa + i);
      if (sb) xb = __ldg(sb + i);
    }
    const float w = win[n] * scale;
    const float2 v = make_float2(xa * w, xb * w);
    e = fmaf(v.x, v.x, fmaf(v.y, v.y, e));
    buf[n] = v;
  }
  return e;
}

// Split the transform Z of (a + i b), a and b real, at bin k (0 <= k <= N/2):
//   A[k] = (Z[k] + conj Z[N-k]) / 2 ,  B[k] = -i (Z[k] - conj Z[N-k]) / 2
__device__ __forceinline__ void unpack_pair(float2 zk, float2 zm, float2& A, float2& B) {
  A = make_float2(0.5f * (zk.x + zm.x), 0.5f * (zk.y - zm.y));
  B = make_float2(0.5f * (zk.y + zm.y), 0.5f * (zm.x - zk.x));  // (zm.x - zk.x): DC/Nyquist give +0, like rfft
}

// ------------------------------------------------------------------------------------------
// TMA bulk copy (cp.async.bulk, SASS UBLKCP) + mbarrier helpers: asynchronous global -> shared staging of
// contiguous 16-byte-aligned blocks, completion signalled as transaction bytes on an mbarrier.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t arrivals) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(arrivals) : "memory");
}
// make barrier initialisation visible to the async proxy
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
// order this thread's generic-proxy accesses to shared memory before later async-proxy (TMA) accesses
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// global -> shared bulk copy of `bytes` (multiple of 16; both addresses 16-byte aligned)
__device__ __forceinline__ void tma_load_1d(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_inval(uint64_t* bar) {
  asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// order earlier generic-proxy writes (of any thread, once observed) before this thread's later async-proxy (TMA) reads
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ int ld_acquire_gpu(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_gpu(int* p, int v) {
  asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// wait until the phase with the given parity has completed; traps instead of hanging if it never does.
// Called by all 32 lanes of a converged warp: the loop exit is a warp vote, so the compiler sees a warp-uniform branch
// and the code behind the wait stays "converged" in its analysis (a per-lane exit made it wrap every later shuffle of
// the frame loop in WARPSYNC.COLLECTIVE - see warp_id_uniform()).
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  for (uint32_t spin = 0;; ++spin) {
    uint32_t done;
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
    if (__all_sync(kFull, done != 0u)) break;
    if (spin > (1u << 22)) {
      printf("mbar_wait timed out (block %d thread %d parity %u)\n", (int)blockIdx.x, (int)threadIdx.x, parity);
      __trap();
    }
  }
}

// atan2 in (-pi, pi] with numpy's conventions for signed zeros: one division, a degree-8 minimax polynomial in
// t^2 for atan(t)/t on [0, 1] (Remez fit, |error| < 1.2e-7 rad evaluated in float32) and quadrant fix-ups - about a
// third of the instructions of atan2f, which dominated the feature kernels.
__device__ __forceinline__ float atan2_poly(float y, float x) {
  const float ax = fabsf(x), ay = fabsf(y);
  const float mx = fmaxf(ax, ay), mn = fminf(ax, ay);
  // mn / mx to ~1 ulp: hardware reciprocal + one Newton step on the quotient (the fast division's 2 ulp would be a
  // quarter of the 1e-6 rad budget of the IPD; a correctly rounded division costs three times as many instructions)
  float t = 0.f;
  if (mx > 0.f) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(mx));
    const float q0 = mn * r;
    t = fmaf(fmaf(-q0, mx, mn), r, q0);
  }
  const float s = t * t;
  float p = 0.0029035559807253364f;
  p = fmaf(p, s, -0.016283021665709413f);
  p = fmaf(p, s, 0.043039389735442454f);
  p = fmaf(p, s, -0.07533677875170354f);
  p = fmaf(p, s, 0.10654678222445768f);
  p = fmaf(p, s, -0.1420713385858043f);
  p = fmaf(p, s, 0.19993054130923057f);
  p = fmaf(p, s, -0.3333309395827876f);
  p = fmaf(p, s, 0.9999999863667985f);
  float r = p * t;
  if (ay > ax) r = 1.57079632679489662f - r;
  if (__float_as_int(x) < 0) r = 3.14159265358979324f - r;   // x < 0 or x == -0
  return copysignf(r, y);
}

// ln(x) for normal x > 0 with an ABSOLUTE error of ~1.3e-7 + half an ulp of the result: x = 2^e f with f in
// [0.7071, 1.4142), ln f from the hardware lg2 (absolute error 2^-22.6 on that interval), e ln 2 added with ln 2 split
// in two so that the product is exact.  9 instructions; logf costs ~40, and __logf (lg2 on the raw argument) carries a
// RELATIVE error of 2^-22 on a result of up to 16, i.e. up to 4e-6 - twice the budget below.
__device__ __forceinline__ float log_abs_accurate(float x) {
  const int ix = __float_as_int(x);
  const int e = (ix - 0x3f3504f3) >> 23;
  const float f = __int_as_float(ix - (e << 23));
  float l2;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l2) : "f"(f));
  const float fe = (float)e;
  return fmaf(fe, 0.693145751953125f, fmaf(fe, 1.42860682030941723e-6f, l2 * 0.693147180559945309f));
}

// log-magnitude + inter-channel phase difference of one TF bin (full_audio.../inference.py:91-94).  The reference
// evaluates np.abs / np.log / np.angle in float64 and casts to float32.  Error budget against that for the SAME
// spectrum (tests/test_gpu_parity.py::test_features_same_spectrum: log-mag <= 2e-6, IPD <= 1e-6 rad on every non-zero
// bin): |y| = p * rsq(p) (2 ulp relative = 2.4e-7 absolute in the logarithm), log_abs_accurate (1.3e-7 + rounding of a
// value of at most 16.2: 0.95e-6), two atan2_poly angles (1.2e-7 polynomial + 0.6e-7 quotient + 1.2e-7 rounding each)
// and the rounding of their difference (2.4e-7).
__device__ __forceinline__ void feature_values(float2 y0, float2 y1, float& logmag, float& ipd) {
  const float p = fmaf(y0.x, y0.x, y0.y * y0.y);
  logmag = log_abs_accurate(p * rsqrtf(fmaxf(p, 1e-37f)) + 1e-7f);   // p = 0 gives |y| = 0, not NaN
  ipd = atan2_poly(y0.y, y0.x) - atan2_poly(y1.y, y1.x);
}

// Store one bin's features in the layout `mode` asks for (AVZ_FEAT_*).
__device__ __forceinline__ void store_features(float* __restrict__ X, int mode, int b, int k, int t, int F, int T,
                                               float lm, float ipd) {
  if (mode == AVZ_FEAT_PHYSICS_NHWC) {
    float s, c;
    sincosf(ipd, &s, &c);
    reinterpret_cast<float4*>(X)[((int64_t)b * F + k) * T + t] =
        make_float4(lm, s, c, (float)((double)k / (double)(F - 1)));
  } else {
    if (mode == AVZ_FEAT_LOGMAG_IPD_WRAPPED) {
      const float two_pi = 6.28318530717958647692f;
      ipd = ipd - two_pi * rintf(ipd / two_pi);  // wrap to [-pi, pi]
    }
    X[(((int64_t)b * 2 + 0) * F + k) * T + t] = lm;
    X[(((int64_t)b * 2 + 1) * F + k) * T + t] = ipd;
  }
}

}  // namespace avz

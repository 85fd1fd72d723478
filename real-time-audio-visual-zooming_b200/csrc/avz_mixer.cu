// On-device far-field 2-mic mixer and the PCM16 wire-format converters (8-F rank 4).
//
// Reference: rt_av_zoom/core/tf_lite_version/world_building.py:46-52 (apply_frac_delay) and :61-93 (mix_and_save).
// The reference delays every source with a whole-signal real FFT of length L (pocketfft handles any L), a phase ramp
// and an inverse real FFT, once per source and microphone.  Here the mixture is formed in the frequency domain:
//
//   pass 1+2  forward length-L FFT of the sources, two real sources packed into one complex transform
//   pass 3    per bin: unpack the sources, apply the 2S phase ramps, sum into mic 1 / mic 2 / target image /
//             interferer image, re-pack as (mic1 + i mic2) and (target + i interferer) Hermitian pairs
//   pass 4+5  inverse length-L FFT of the two packed outputs, 1/L, running max|mix|
//   pass 6    divide everything by max|mix| + eps
//
// so S sources cost ceil(S/2) + 2 complex transforms instead of the reference's 3S real ones.
//
// The length-L transform is a two-factor Cooley-Tukey split L = N1*N2 (N1 = 2^a <= 512, N2 <= 1024 arbitrary):
// with n = N2 n1 + n2 and k = k1 + N1 k2,
//   X[k1 + N1 k2] = sum_n2 W_N2^{n2 k2} * [ W_L^{n2 k1} * sum_n1 x[N2 n1 + n2] W_N1^{n1 k1} ]
// "cols" kernels do the power-of-two part over n1 (16 adjacent n2 columns per block so that global accesses are
// 64/128-byte runs; radix-2 in shared memory, DIF forward / DIT inverse so that no bit-reversal pass is needed),
// "rows" kernels do the length-N2 part as a direct DFT on contiguous rows (N2 = 125 for 4 s at 16 kHz).
// Spectra stay in the [k1][k2] order between the passes; the inverse runs the two factors in the opposite order and
// lands in natural time order, so no transposition is ever materialised.  All twiddles come from one table
// W_L[j] = exp(-2 pi i j / L) rounded from float64 (W_N1^j = W_L[j N2], W_N2^j = W_L[j N1]).
#include <cmath>
#include <cstdlib>
#include <mutex>
#include <vector>

#include "avz_common.cuh"

namespace avz {
namespace {

constexpr int kMaxSrc = 8;
constexpr int kCols = 16;        // n2 columns per block in the power-of-two passes
constexpr int kColThreads = 1024;  // 2 CTAs of 64 KB per SM -> 64 warps at 32 registers (256 threads: 24 warps, 13 % slower end to end)
constexpr int kRowsPer = 4;      // rows sharing one twiddle fetch in the direct-DFT passes
constexpr int kMaxN1 = 512;
constexpr int kMaxN2 = 1024;

struct MixTable {
  int device;
  int64_t n;
  const float2* w;
};
std::mutex g_mix_mu;
std::vector<MixTable> g_mix_tables;

int mix_table_for(int64_t n, const float2** out) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return set_error(AVZ_ENOGPU, "cudaGetDevice: %s", cudaGetErrorString(e));
  std::lock_guard<std::mutex> lk(g_mix_mu);
  for (const auto& t : g_mix_tables)
    if (t.device == dev && t.n == n) {
      *out = t.w;
      return AVZ_OK;
    }
  std::vector<float2> w((size_t)n);
  const double two_pi = 6.283185307179586476925286766559;
  for (int64_t j = 0; j < n; ++j) {
    double c = cos(two_pi * (double)j / (double)n), s = -sin(two_pi * (double)j / (double)n);
    if ((4 * j) % n == 0) {  // exact quadrant points
      const int q = (int)((4 * j) / n);
      c = (q == 0) ? 1.0 : (q == 2 ? -1.0 : 0.0);
      s = (q == 1) ? -1.0 : (q == 3 ? 1.0 : 0.0);
    }
    w[(size_t)j] = make_float2((float)c, (float)s);
  }
  void* p = nullptr;
  AVZ_CUDA_OK(cudaMalloc(&p, (size_t)n * sizeof(float2)));
  AVZ_CUDA_OK(cudaMemcpy(p, w.data(), (size_t)n * sizeof(float2), cudaMemcpyHostToDevice));
  g_mix_tables.push_back({dev, n, (const float2*)p});
  *out = (const float2*)p;
  return AVZ_OK;
}

// L = N1 * N2: N1 = largest power of two <= kMaxN1 dividing L.
bool split_length(int64_t L, int* N1, int* log2N1, int* N2) {
  if (L < 2) return false;
  int n1 = 1, lg = 0;
  while (n1 < kMaxN1 && L % (2 * (int64_t)n1) == 0) { n1 *= 2; ++lg; }
  const int64_t n2 = L / n1;
  if (n2 > kMaxN2) return false;
  *N1 = n1;
  *log2N1 = lg;
  *N2 = (int)n2;
  return true;
}

// The power-of-two passes run radix-4 stages (quarter sizes N1/4, N1/16, ...) plus one radix-2 stage when log2 N1 is
// odd, in place: decimation in frequency leaves bin k at the position whose base-4 digits (and last bit) are k's in
// reverse order; the inverse places bin k there and runs the stages backwards (decimation in time).
__device__ __forceinline__ int pos_to_bin(int pos, int N1) {
  int k = 0, mult = 1, rem = N1, p = pos;
  while (rem >= 4) {
    const int q = rem >> 2, d = p / q;
    p -= d * q;
    k += d * mult;
    mult <<= 2;
    rem = q;
  }
  if (rem == 2) k += p * mult;
  return k;
}
__device__ __forceinline__ int bin_to_pos(int k, int N1) {
  int pos = 0, rem = N1;
  while (rem >= 4) {
    const int q = rem >> 2;
    pos += (k & 3) * q;
    k >>= 2;
    rem = q;
  }
  if (rem == 2) pos += k & 1;
  return pos;
}

// What the forward power-of-two pass reads (the Bluestein variants serve lengths that have no N1*N2 split, see below).
enum ColLoad {
  kLoadPair = 0,      // two real sources of length N packed as re + i im
  kLoadBluePair = 1,  // the same, of length Ls <= N: times the chirp c[n], zero beyond Ls
  kLoadBlueSpec = 2,  // a complex spectrum of length Ls: conj(Y[k]) * c[k], zero beyond Ls (inverse transform)
  kLoadPlane = 3      // a complex plane of length N as it is
};
// What the inverse power-of-two pass writes.
enum ColStore {
  kStoreReal = 0,      // re / im of x / N into two real signals of length N (+ running max|mix|)
  kStoreBlueSpec = 1,  // conv[k] * c[k] / N for k < Ls into a complex spectrum of length Ls (natural bin order)
  kStoreBlueReal = 2   // conj(conv[n] * c[n]) / (N Ls) for n < Ls into two real signals of length Ls (+ max|mix|)
};

// ---- pass 1: pack two sources, power-of-two DIF over n1, twiddle W_L^{n2 k1}; A[b][p][k1][n2] ----
//      N = N1*N2 is the transform length (= plane stride of A); Ls the signal length (= N unless Bluestein).
template <int LOAD>
__global__ void __launch_bounds__(kColThreads, 2)
k_mix_cols_fwd(const float* __restrict__ src, const float2* __restrict__ spec, const float2* __restrict__ chirp,
               float2* __restrict__ A, const float2* __restrict__ W, int S, int PP, int N1, int log2N1, int N2,
               int64_t N, int64_t Ls) {
  extern __shared__ float2 sm[];  // [N1][kCols]
  const int c0 = blockIdx.x * kCols;
  const int p = blockIdx.y, b = blockIdx.z;
  const int ncol = min(kCols, N2 - c0);
  const float* sa = src + ((int64_t)b * S + 2 * p) * Ls;
  const bool has_b = 2 * p + 1 < S;
  const float* sb = sa + Ls;
  const float2* sp = spec + ((int64_t)b * PP + p) * Ls;
  for (int idx = threadIdx.x; idx < N1 * kCols; idx += kColThreads) {
    const int c = idx & (kCols - 1), n1 = idx / kCols;
    float2 z = make_float2(0.f, 0.f);
    const int64_t g = (int64_t)n1 * N2 + c0 + c;
    if (c < ncol && g < Ls) {
      if (LOAD == kLoadPair || LOAD == kLoadBluePair) {
        z.x = __ldg(sa + g);
        if (has_b) z.y = __ldg(sb + g);
        if (LOAD == kLoadBluePair) z = cmul(z, __ldg(chirp + g));
      } else if (LOAD == kLoadBlueSpec) {
        const float2 y = __ldg(sp + g);
        z = cmul(make_float2(y.x, -y.y), __ldg(chirp + g));
      } else {
        z = __ldg(sp + g);
      }
    }
    sm[idx] = z;
  }
  __syncthreads();
  int rem = N1;
  for (; rem >= 4; rem >>= 2) {   // radix-4 DIF stage on blocks of `rem`
    const int q = rem >> 2;
    const int64_t tstep = (int64_t)N2 * (N1 / rem);  // W_rem^j = W_L[j * L / rem]
    for (int idx = threadIdx.x; idx < (N1 / 4) * kCols; idx += kColThreads) {
      const int c = idx & (kCols - 1), j = idx / kCols;
      const int pos = j & (q - 1);
      const int i0 = ((j - pos) << 2) + pos;
      float2* x = sm + i0 * kCols + c;
      const float2 a0 = x[0], a1 = x[q * kCols], a2 = x[2 * q * kCols], a3 = x[3 * q * kCols];
      const float2 t0 = cadd(a0, a2), t1 = csub(a0, a2), t2 = cadd(a1, a3);
      const float2 d = csub(a1, a3);
      const float2 t3 = make_float2(d.y, -d.x);   // * (-i)
      x[0] = cadd(t0, t2);
      x[q * kCols] = cmul(cadd(t1, t3), __ldg(W + pos * tstep));
      x[2 * q * kCols] = cmul(csub(t0, t2), __ldg(W + 2 * pos * tstep));
      x[3 * q * kCols] = cmul(csub(t1, t3), __ldg(W + 3 * pos * tstep));
    }
    __syncthreads();
  }
  if (rem == 2) {
    for (int idx = threadIdx.x; idx < (N1 / 2) * kCols; idx += kColThreads) {
      const int c = idx & (kCols - 1), j = idx / kCols;
      float2* x = sm + 2 * j * kCols + c;
      const float2 a = x[0], bb = x[kCols];
      x[0] = cadd(a, bb);
      x[kCols] = csub(a, bb);
    }
    __syncthreads();
  }
  // position -> bin once per row of the tile (the digit-reversal loop with its divisions used to run per element)
  __shared__ unsigned short s_bin[kMaxN1];
  for (int pos = threadIdx.x; pos < N1; pos += kColThreads) s_bin[pos] = (unsigned short)pos_to_bin(pos, N1);
  __syncthreads();
  float2* out = A + ((int64_t)b * PP + p) * N;
  for (int idx = threadIdx.x; idx < N1 * kCols; idx += kColThreads) {
    const int c = idx & (kCols - 1), pos = idx / kCols;
    if (c < ncol) {
      const int k1 = s_bin[pos];
      const int n2 = c0 + c;
      out[(int64_t)k1 * N2 + n2] = cmul(sm[idx], __ldg(W + (int64_t)n2 * k1));
    }
  }
}

// ---- passes 2 and 4: length-N2 DFT along contiguous rows, in place.  N2 = Na*Nb (Na the largest divisor <= sqrt N2,
//      1 for a prime N2) is split once more inside shared memory: with n = Nb na + nb and k = ka + Na kb,
//        t[ka][nb] = W_N2^{nb ka} * sum_na x[Nb na + nb] W_Na^{na ka}        (Na terms per output)
//        X[ka + Na kb] = sum_nb t[ka][nb] W_Nb^{nb kb}                         (Nb terms per output)
//      i.e. N2*(Na+Nb) instead of N2^2 complex multiply-adds per row (125 = 5*25: 4x fewer).  One thread per output
//      index, kRowsPer rows share every twiddle fetch; wm[j] = W_N2^j (conjugated for the inverse).
//      `mul` (optional, [k1][k2] like the data): every element is multiplied by it on the way in - the Bluestein
//      convolution kernel's spectrum, applied where the inverse transform starts. ----
template <bool INV>
__global__ void k_mix_rows(float2* __restrict__ Z, const float2* __restrict__ W, const float2* __restrict__ mul, int N1,
                           int N2, int Na, int Nb, int64_t N) {
  extern __shared__ float2 sm[];  // wm[N2] | rows[kRowsPer][N2] | tb[kRowsPer][N2]
  float2* wm = sm;
  float2* rows = sm + N2;
  float2* tb = rows + kRowsPer * N2;
  for (int j = threadIdx.x; j < N2; j += blockDim.x) {
    float2 w = __ldg(W + (int64_t)j * N1);
    if (INV) w.y = -w.y;
    wm[j] = w;
  }
  float2* base = Z + (int64_t)blockIdx.y * N;
  const int t = threadIdx.x;
  const int ka = t / Nb, r_ = t - ka * Nb;  // (ka, nb) in step A, (ka, kb) in step B
  for (int r0 = blockIdx.x * kRowsPer; r0 < N1; r0 += gridDim.x * kRowsPer) {
    const int nr = min(kRowsPer, N1 - r0);
    __syncthreads();
    for (int j = threadIdx.x; j < kRowsPer * N2; j += blockDim.x) {
      float2 v = make_float2(0.f, 0.f);
      if (j < nr * N2) {
        v = base[(int64_t)r0 * N2 + j];
        if (mul) v = cmul(v, __ldg(mul + (int64_t)r0 * N2 + j));
      }
      rows[j] = v;
    }
    __syncthreads();
    if (t < N2) {  // step A
      float2 acc[kRowsPer];
#pragma unroll
      for (int r = 0; r < kRowsPer; ++r) acc[r] = make_float2(0.f, 0.f);
      int ia = 0;  // (na * ka) mod Na
      for (int na = 0; na < Na; ++na) {
        const float2 w = wm[ia * Nb];
#pragma unroll
        for (int r = 0; r < kRowsPer; ++r) {
          const float2 v = rows[r * N2 + Nb * na + r_];
          acc[r].x = fmaf(v.x, w.x, fmaf(-v.y, w.y, acc[r].x));
          acc[r].y = fmaf(v.x, w.y, fmaf(v.y, w.x, acc[r].y));
        }
        ia += ka;
        if (ia >= Na) ia -= Na;
      }
      const float2 w2 = wm[r_ * ka];  // nb * ka < N2
#pragma unroll
      for (int r = 0; r < kRowsPer; ++r) tb[r * N2 + t] = cmul(acc[r], w2);
    }
    __syncthreads();
    if (t < N2) {  // step B
      float2 acc[kRowsPer];
#pragma unroll
      for (int r = 0; r < kRowsPer; ++r) acc[r] = make_float2(0.f, 0.f);
      int ib = 0;  // (nb * kb) mod Nb
      for (int nb = 0; nb < Nb; ++nb) {
        const float2 w = wm[ib * Na];
#pragma unroll
        for (int r = 0; r < kRowsPer; ++r) {
          const float2 v = tb[r * N2 + ka * Nb + nb];
          acc[r].x = fmaf(v.x, w.x, fmaf(-v.y, w.y, acc[r].x));
          acc[r].y = fmaf(v.x, w.y, fmaf(v.y, w.x, acc[r].y));
        }
        ib += r_;
        if (ib >= Nb) ib -= Nb;
      }
#pragma unroll
      for (int r = 0; r < kRowsPer; ++r) rows[r * N2 + ka + Na * r_] = acc[r];
    }
    __syncthreads();
    for (int j = threadIdx.x; j < nr * N2; j += blockDim.x) base[(int64_t)r0 * N2 + j] = rows[j];
  }
}

// ---- passes 2 and 4, fast variant: when N2 has only prime factors <= 7 the row transform runs as a mixed-radix
//      Stockham FFT in shared memory (radices <= 8, natural order in and out) instead of the two-level direct DFT:
//      sum(radix) instead of Na + Nb complex multiply-adds per element (125 = 5*5*5: 15 instead of 30) and - what
//      matters more, the direct DFT is shared-memory-bandwidth-bound - a butterfly keeps its R inputs of kRowsPer rows
//      and the R roots of unity in registers: ~0.4 shared-memory loads per multiply-add instead of 1.25. ----
constexpr int kMaxStages = 12;
constexpr int kFftRows = 2;        // rows sharing the twiddles of a butterfly thread
constexpr int kFftThreads = 512;
struct RowFftPlan {
  int n_stages;
  int radix[kMaxStages];
};

template <int R>
__device__ __forceinline__ void row_stage(const float2* __restrict__ in, float2* __restrict__ out,
                                          const float2* __restrict__ wm, int N2, int Ns, int groups) {
  const int nb = N2 / R;                 // butterflies per row
  const int tstep = N2 / (Ns * R);       // W_{Ns R}^{k t} = wm[k t tstep]
  float2 root[R];                        // W_R^m = wm[m nb]
#pragma unroll
  for (int m = 0; m < R; ++m) root[m] = wm[m * nb];
  for (int idx = threadIdx.x; idx < groups * nb; idx += blockDim.x) {
    const int g = idx / nb, j = idx - g * nb;
    const int k = j % Ns;
    const float2* src = in + (size_t)g * kFftRows * N2 + j;
    float2 v[kFftRows][R];
#pragma unroll
    for (int t = 0; t < R; ++t) {
      if (t == 0 || Ns == 1) {
#pragma unroll
        for (int q = 0; q < kFftRows; ++q) v[q][t] = src[q * N2 + t * nb];
      } else {
        const float2 w = wm[k * t * tstep];
#pragma unroll
        for (int q = 0; q < kFftRows; ++q) v[q][t] = cmul(src[q * N2 + t * nb], w);
      }
    }
    float2* dst = out + (size_t)g * kFftRows * N2 + (j - k) * R + k;
#pragma unroll
    for (int u = 0; u < R; ++u) {
#pragma unroll
      for (int q = 0; q < kFftRows; ++q) {
        float2 a = v[q][0];
#pragma unroll
        for (int t = 1; t < R; ++t) {
          const float2 w = root[(t * u) % R];
          const float2 x = v[q][t];
          a.x = fmaf(x.x, w.x, fmaf(-x.y, w.y, a.x));
          a.y = fmaf(x.x, w.y, fmaf(x.y, w.x, a.y));
        }
        dst[q * N2 + u * Ns] = a;
      }
    }
  }
}

template <bool INV>
__global__ void __launch_bounds__(kFftThreads, 3)
k_mix_rows_fft(float2* __restrict__ Z, const float2* __restrict__ W, const float2* __restrict__ mul, int N1, int N2,
               int groups, RowFftPlan plan, int64_t N) {
  extern __shared__ float2 sm[];  // wm[N2] | bufA[groups*kFftRows][N2] | bufB[same]
  float2* wm = sm;
  float2* bufA = sm + N2;
  float2* bufB = bufA + (size_t)groups * kFftRows * N2;
  for (int j = threadIdx.x; j < N2; j += blockDim.x) {
    float2 w = __ldg(W + (int64_t)j * N1);
    if (INV) w.y = -w.y;
    wm[j] = w;
  }
  float2* base = Z + (int64_t)blockIdx.y * N;
  const int rows_blk = groups * kFftRows;
  for (int r0 = blockIdx.x * rows_blk; r0 < N1; r0 += gridDim.x * rows_blk) {
    const int nr = min(rows_blk, N1 - r0);
    __syncthreads();
    for (int j = threadIdx.x; j < rows_blk * N2; j += blockDim.x) {
      float2 v = make_float2(0.f, 0.f);
      if (j < nr * N2) {
        v = base[(int64_t)r0 * N2 + j];
        if (mul) v = cmul(v, __ldg(mul + (int64_t)r0 * N2 + j));
      }
      bufA[j] = v;
    }
    __syncthreads();
    float2* in = bufA;
    float2* out = bufB;
    int Ns = 1;
    for (int s = 0; s < plan.n_stages; ++s) {
      switch (plan.radix[s]) {
        case 2: row_stage<2>(in, out, wm, N2, Ns, groups); break;
        case 3: row_stage<3>(in, out, wm, N2, Ns, groups); break;
        case 4: row_stage<4>(in, out, wm, N2, Ns, groups); break;
        case 5: row_stage<5>(in, out, wm, N2, Ns, groups); break;
        case 7: row_stage<7>(in, out, wm, N2, Ns, groups); break;
        default: row_stage<8>(in, out, wm, N2, Ns, groups); break;
      }
      Ns *= plan.radix[s];
      __syncthreads();
      float2* tmp = in; in = out; out = tmp;
    }
    for (int j = threadIdx.x; j < nr * N2; j += blockDim.x) base[(int64_t)r0 * N2 + j] = in[j];
  }
}

struct MixParams {
  int S;
  int sym[kMaxSrc];    // c2 == -c1 (the two microphones sit symmetrically about the array centre): ramp 2 = conj(ramp 1)
  double c1[kMaxSrc];  // tau(s, mic 1) * fs / L  (cycles per bin)
  double c2[kMaxSrc];
};

__device__ __forceinline__ float2 ramp(double cyc_per_bin, int k) {
  float s, c;
  sincospif((float)(2.0 * (double)k * cyc_per_bin), &s, &c);
  return make_float2(c, -s);  // exp(-2 pi i k c)
}

// ---- pass 3: unpack sources, phase ramps, sum, re-pack.  In place on the spectra: the thread that owns the pair
//      (k, L-k) is the only one that touches those two positions of any plane. ----
__global__ void k_mix_combine(float2* __restrict__ Z, MixParams prm, int PP, int N1, int N2, int64_t N) {
  // 32-bit index arithmetic (N <= 512 * 1024): the 64-bit divisions of the first version were a third of the kernel
  const int n = (int)N;
  const int o = blockIdx.x * blockDim.x + threadIdx.x;
  if (o >= n) return;
  const int b = blockIdx.y;
  const int k1 = o / N2, k2 = o - k1 * N2;
  const int k = k1 + N1 * k2;
  const int km = (k == 0) ? 0 : n - k;
  if (k > km) return;
  const int kq = km / N1;
  const int om = (km - kq * N1) * N2 + kq;
  float2* zb = Z + (int64_t)b * PP * N;
  float2 m1 = make_float2(0.f, 0.f), m2 = m1, tg = m1;
  const int P = (prm.S + 1) / 2;
  for (int p = 0; p < P; ++p) {
    const float2 zk = zb[(int64_t)p * N + o];
    const float2 zm = zb[(int64_t)p * N + om];
    const float2 a = make_float2(0.5f * (zk.x + zm.x), 0.5f * (zk.y - zm.y));
    const float2 bv = make_float2(0.5f * (zk.y + zm.y), 0.5f * (zm.x - zk.x));
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int s = 2 * p + h;
      if (s < prm.S) {
        const float2 v = h ? bv : a;
        const float2 r1 = ramp(prm.c1[s], k);
        const float2 r2 = prm.sym[s] ? make_float2(r1.x, -r1.y) : ramp(prm.c2[s], k);   // sincospif is exactly odd / even
        const float2 d1 = cmul(v, r1);
        const float2 d2 = cmul(v, r2);
        m1 = cadd(m1, d1);
        m2 = cadd(m2, d2);
        if (s == 0) tg = d1;
      }
    }
  }
  float2 in = csub(m1, tg);
  if (k == km) {  // DC and (even L) Nyquist: the real inverse transform ignores the imaginary part
    m1.y = 0.f; m2.y = 0.f; tg.y = 0.f; in.y = 0.f;
  }
  // x + i y for two real signals x, y: bin k holds X + iY, bin L-k holds conj(X) + i conj(Y)
  zb[o] = make_float2(m1.x - m2.y, m1.y + m2.x);
  zb[N + o] = make_float2(tg.x - in.y, tg.y + in.x);
  if (om != o) {
    zb[om] = make_float2(m1.x + m2.y, m2.x - m1.y);
    zb[N + om] = make_float2(tg.x + in.y, in.x - tg.y);
  }
}

// ---- pass 5: twiddle conj(W_L^{n2 k1}), power-of-two DIT over k1, 1/L, write the four real signals ----
template <int STORE>
__global__ void __launch_bounds__(kColThreads, 2)
k_mix_cols_inv(const float2* __restrict__ Q, const float2* __restrict__ W, const float2* __restrict__ chirp,
               float2* __restrict__ spec, float* __restrict__ mix, float* __restrict__ tgt, float* __restrict__ itf,
               unsigned* __restrict__ peak_bits, int PP, int N1, int log2N1, int N2, int64_t N, int64_t Ls) {
  extern __shared__ float2 sm[];
  __shared__ float s_max[kColThreads / 32];
  const int c0 = blockIdx.x * kCols;
  const int q = blockIdx.y, b = blockIdx.z;
  const int ncol = min(kCols, N2 - c0);
  const float2* in = Q + ((int64_t)b * PP + q) * N;
  for (int idx = threadIdx.x; idx < N1 * kCols; idx += kColThreads) {
    const int c = idx & (kCols - 1), k1 = idx / kCols;
    float2 z = make_float2(0.f, 0.f);
    if (c < ncol) {
      const int n2 = c0 + c;
      z = cmulc(in[(int64_t)k1 * N2 + n2], __ldg(W + (int64_t)n2 * k1));
    }
    sm[bin_to_pos(k1, N1) * kCols + c] = z;   // shifts and masks only (a lookup table here measured slower)
  }
  __syncthreads();
  int rem = (log2N1 & 1) ? 2 : 4;
  if (rem == 2 && N1 >= 2) {
    for (int idx = threadIdx.x; idx < (N1 / 2) * kCols; idx += kColThreads) {
      const int c = idx & (kCols - 1), j = idx / kCols;
      float2* x = sm + 2 * j * kCols + c;
      const float2 a = x[0], bb = x[kCols];
      x[0] = cadd(a, bb);
      x[kCols] = csub(a, bb);
    }
    __syncthreads();
    rem = 8;
  }
  for (; rem <= N1; rem <<= 2) {   // radix-4 DIT stage on blocks of `rem`
    const int q = rem >> 2;
    const int64_t tstep = (int64_t)N2 * (N1 / rem);
    for (int idx = threadIdx.x; idx < (N1 / 4) * kCols; idx += kColThreads) {
      const int c = idx & (kCols - 1), j = idx / kCols;
      const int pos = j & (q - 1);
      const int i0 = ((j - pos) << 2) + pos;
      float2* x = sm + i0 * kCols + c;
      const float2 b0 = x[0];
      const float2 b1 = cmulc(x[q * kCols], __ldg(W + pos * tstep));
      const float2 b2 = cmulc(x[2 * q * kCols], __ldg(W + 2 * pos * tstep));
      const float2 b3 = cmulc(x[3 * q * kCols], __ldg(W + 3 * pos * tstep));
      const float2 t0 = cadd(b0, b2), t1 = csub(b0, b2), t2 = cadd(b1, b3);
      const float2 d = csub(b1, b3);
      const float2 t3 = make_float2(-d.y, d.x);   // * (+i)
      x[0] = cadd(t0, t2);
      x[q * kCols] = cadd(t1, t3);
      x[2 * q * kCols] = csub(t0, t2);
      x[3 * q * kCols] = csub(t1, t3);
    }
    __syncthreads();
  }
  if (STORE == kStoreBlueSpec) {
    const float inv_n = (float)(1.0 / (double)N);
    float2* o = spec + ((int64_t)b * PP + q) * Ls;
    for (int idx = threadIdx.x; idx < N1 * kCols; idx += kColThreads) {
      const int c = idx & (kCols - 1), n1 = idx / kCols;
      const int64_t g = (int64_t)n1 * N2 + c0 + c;
      if (c < ncol && g < Ls) {
        const float2 v = cmul(sm[idx], __ldg(chirp + g));
        o[g] = make_float2(v.x * inv_n, v.y * inv_n);
      }
    }
    return;
  }
  const float inv_n = (float)(STORE == kStoreBlueReal ? 1.0 / ((double)N * (double)Ls) : 1.0 / (double)N);
  float* o_re = q == 0 ? mix + (int64_t)b * 2 * Ls : tgt + (int64_t)b * Ls;
  float* o_im = q == 0 ? mix + ((int64_t)b * 2 + 1) * Ls : itf + (int64_t)b * Ls;
  float mx = 0.f;
  for (int idx = threadIdx.x; idx < N1 * kCols; idx += kColThreads) {
    const int c = idx & (kCols - 1), n1 = idx / kCols;
    const int64_t g = (int64_t)n1 * N2 + c0 + c;
    if (c < ncol && g < Ls) {
      float2 v = sm[idx];
      if (STORE == kStoreBlueReal) {
        v = cmul(v, __ldg(chirp + g));
        v.y = -v.y;
      }
      const float re = v.x * inv_n, im = v.y * inv_n;
      o_re[g] = re;
      o_im[g] = im;
      mx = fmaxf(mx, fmaxf(fabsf(re), fabsf(im)));
    }
  }
  if (q == 0) {  // uniform per block
    mx = warp_max(mx);
    if ((threadIdx.x & 31) == 0) s_max[threadIdx.x >> 5] = mx;
    __syncthreads();
    if (threadIdx.x == 0) {
      for (int w = 1; w < kColThreads / 32; ++w) mx = fmaxf(mx, s_max[w]);
      atomicMax(peak_bits + b, __float_as_uint(mx));  // non-negative floats order like their bit patterns
    }
  }
}

// ---- pass 6: divide mix / tgt / itf by max|mix| + eps (world_building.py:86-91) ----
__global__ void __launch_bounds__(256)
k_mix_scale(float* __restrict__ mix, float* __restrict__ tgt, float* __restrict__ itf,
            const unsigned* __restrict__ peak_bits, float peak_eps, int64_t N) {
  const int b = blockIdx.y;
  const float den = __uint_as_float(peak_bits[b]) + peak_eps;
  // the four signals of an utterance, float4 at a time when N and the caller's pointers allow it (each signal then
  // starts 16-byte aligned); a shifted view of a larger buffer takes the scalar loop
  if ((N & 3) == 0 && ((reinterpret_cast<uintptr_t>(mix) | reinterpret_cast<uintptr_t>(tgt) | reinterpret_cast<uintptr_t>(itf)) & 15) == 0) {
    const int64_t n4 = N >> 2;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < 4 * n4; i += (int64_t)gridDim.x * blockDim.x) {
      float4* p = i < 2 * n4 ? reinterpret_cast<float4*>(mix + (int64_t)b * 2 * N) + i
                             : (i < 3 * n4 ? reinterpret_cast<float4*>(tgt + (int64_t)b * N) + (i - 2 * n4)
                                           : reinterpret_cast<float4*>(itf + (int64_t)b * N) + (i - 3 * n4));
      const float4 v = __ldcs(p);
      *p = make_float4(__fdiv_rn(v.x, den), __fdiv_rn(v.y, den), __fdiv_rn(v.z, den), __fdiv_rn(v.w, den));
    }
    return;
  }
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < 4 * N; i += (int64_t)gridDim.x * blockDim.x) {
    float* p = i < 2 * N ? mix + (int64_t)b * 2 * N + i : (i < 3 * N ? tgt + (int64_t)b * N + (i - 2 * N)
                                                                      : itf + (int64_t)b * N + (i - 3 * N));
    *p = __fdiv_rn(*p, den);
  }
}

__global__ void k_pcm16_to_f32(const int16_t* __restrict__ pcm, int64_t n, float* __restrict__ out) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t n8 = n / 8;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += stride) {
    const int4 raw = __ldcs(reinterpret_cast<const int4*>(pcm) + i);
    const int w[4] = {raw.x, raw.y, raw.z, raw.w};
    float4 lo, hi;
    lo.x = (float)(short)(w[0] & 0xffff) * (1.f / 32768.f);
    lo.y = (float)(short)(w[0] >> 16) * (1.f / 32768.f);
    lo.z = (float)(short)(w[1] & 0xffff) * (1.f / 32768.f);
    lo.w = (float)(short)(w[1] >> 16) * (1.f / 32768.f);
    hi.x = (float)(short)(w[2] & 0xffff) * (1.f / 32768.f);
    hi.y = (float)(short)(w[2] >> 16) * (1.f / 32768.f);
    hi.z = (float)(short)(w[3] & 0xffff) * (1.f / 32768.f);
    hi.w = (float)(short)(w[3] >> 16) * (1.f / 32768.f);
    reinterpret_cast<float4*>(out)[2 * i] = lo;
    reinterpret_cast<float4*>(out)[2 * i + 1] = hi;
  }
  for (int64_t i = n8 * 8 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    out[i] = (float)pcm[i] * (1.f / 32768.f);
}

__device__ __forceinline__ int to_pcm(float x) {
  // libsndfile f2s: lrintf(x * 0x7FFF); NaN -> 0; clipped instead of wrapping
  const float v = x * 32767.f;
  if (!(v == v)) return 0;
  return max(-32768, min(32767, __float2int_rn(fminf(fmaxf(v, -40000.f), 40000.f))));
}

__global__ void k_f32_to_pcm16(const float* __restrict__ x, int64_t n, int16_t* __restrict__ pcm) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t n8 = n / 8;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += stride) {
    const float4 lo = __ldcs(reinterpret_cast<const float4*>(x) + 2 * i);
    const float4 hi = __ldcs(reinterpret_cast<const float4*>(x) + 2 * i + 1);
    int4 o;
    o.x = (int)((unsigned)(to_pcm(lo.x) & 0xffff) | ((unsigned)to_pcm(lo.y) << 16));
    o.y = (int)((unsigned)(to_pcm(lo.z) & 0xffff) | ((unsigned)to_pcm(lo.w) << 16));
    o.z = (int)((unsigned)(to_pcm(hi.x) & 0xffff) | ((unsigned)to_pcm(hi.y) << 16));
    o.w = (int)((unsigned)(to_pcm(hi.z) & 0xffff) | ((unsigned)to_pcm(hi.w) << 16));
    reinterpret_cast<int4*>(pcm)[i] = o;
  }
  for (int64_t i = n8 * 8 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    pcm[i] = (int16_t)to_pcm(x[i]);
}


// ---- host side: transform geometry, launch helpers, chirp-z plans ----
struct Split {
  int N1 = 1, lg = 0, N2 = 1;
  int64_t len() const { return (int64_t)N1 * N2; }
  size_t col_smem() const { return (size_t)N1 * kCols * sizeof(float2); }
  size_t row_smem() const { return (size_t)(1 + 2 * kRowsPer) * N2 * sizeof(float2); }
  void row_factors(int* Na, int* Nb) const {
    int a = 1;
    for (int d = 1; d * d <= N2; ++d)
      if (N2 % d == 0) a = d;
    *Na = a;
    *Nb = N2 / a;
  }
};

// Convolution length for the chirp-z path: M = 2^a * N2 >= 2L-1 with a <= 9, 2 <= N2 <= kMaxN2, chosen to minimise a
// simple cost model (M * (2 a + Na + Nb): radix stages of the column pass + terms of the two-level row DFT).
bool bluestein_split(int64_t L, Split* out) {
  if (L < 1) return false;
  const int64_t need = 2 * L - 1;
  if (need > (int64_t)kMaxN1 * kMaxN2) return false;
  double best = 0.0;
  bool found = false;
  for (int a = 0; (1 << a) <= kMaxN1; ++a) {
    const int64_t n1 = (int64_t)1 << a;
    int64_t lo = (need + n1 - 1) / n1;
    if (lo < 2) lo = 2;
    for (int64_t n2 = lo; n2 <= kMaxN2 && n2 < lo + 48; ++n2) {
      Split c;
      c.N1 = (int)n1;
      c.lg = a;
      c.N2 = (int)n2;
      int Na, Nb;
      c.row_factors(&Na, &Nb);
      const double cost = (double)c.len() * (2.0 * a + Na + Nb);
      if (!found || cost < best) {
        best = cost;
        *out = c;
        found = true;
      }
    }
  }
  return found;
}

int set_mix_attrs(const Split& sp) {
  static std::mutex attr_mu;
  std::lock_guard<std::mutex> lk(attr_mu);
  const int cs = (int)sp.col_smem(), rs = (int)sp.row_smem();
  AVZ_CUDA_OK(cudaFuncSetAttribute(k_mix_cols_fwd<kLoadPair>, cudaFuncAttributeMaxDynamicSharedMemorySize, cs));
  AVZ_CUDA_OK(cudaFuncSetAttribute(k_mix_cols_fwd<kLoadBluePair>, cudaFuncAttributeMaxDynamicSharedMemorySize, cs));
  AVZ_CUDA_OK(cudaFuncSetAttribute(k_mix_cols_fwd<kLoadBlueSpec>, cudaFuncAttributeMaxDynamicSharedMemorySize, cs));
  AVZ_CUDA_OK(cudaFuncSetAttribute(k_mix_cols_fwd<kLoadPlane>, cudaFuncAttributeMaxDynamicSharedMemorySize, cs));
  AVZ_CUDA_OK(cudaFuncSetAttribute(k_mix_cols_inv<kStoreReal>, cudaFuncAttributeMaxDynamicSharedMemorySize, cs));
  AVZ_CUDA_OK(cudaFuncSetAttribute(k_mix_cols_inv<kStoreBlueSpec>, cudaFuncAttributeMaxDynamicSharedMemorySize, cs));
  AVZ_CUDA_OK(cudaFuncSetAttribute(k_mix_cols_inv<kStoreBlueReal>, cudaFuncAttributeMaxDynamicSharedMemorySize, cs));
  AVZ_CUDA_OK(cudaFuncSetAttribute(k_mix_rows<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, rs));
  AVZ_CUDA_OK(cudaFuncSetAttribute(k_mix_rows<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, rs));
  return AVZ_OK;
}

// Radices (<= 8) of the mixed-radix row transform, or false when N2 has a prime factor above 7.
bool row_fft_plan(int N2, RowFftPlan* plan) {
  int n = N2, ns = 0;
  const int order[] = {5, 7, 3, 8, 4, 2};
  for (int r : order)
    while (n % r == 0 && n > 1) {
      if (ns == kMaxStages) return false;
      plan->radix[ns++] = r;
      n /= r;
    }
  plan->n_stages = ns;
  return n == 1 && ns > 0;
}

// Row pass over planes 0 .. used-1 of every utterance (PP planes of sp.len() elements per utterance).
template <bool INV>
int launch_rows(float2* Z, const float2* W, const float2* mul, const Split& sp, int B, int used, int PP, cudaStream_t st) {
  if (sp.N2 <= 1) return AVZ_OK;
  RowFftPlan plan;
#ifdef AVZ_EXPERIMENT
  static const bool direct_only = [] {
    const char* e = getenv("AVZ_MIXER_DIRECT_ROWS");   // A/B: force the two-level direct DFT
    return e && e[0] == '1';
  }();
#else
  constexpr bool direct_only = false;
#endif
  if (!direct_only && row_fft_plan(sp.N2, &plan)) {
    const int64_t M = sp.len();
    // 32 rows per block when they fit in ~72 KB (3 blocks of 512 threads per SM = 48 warps: the pass is latency-bound,
    // occupancy is what moves it - 256 threads x 4 blocks measured 6 % slower end to end), fewer for long rows
    int groups = (int)((72 * 1024 / (int64_t)sizeof(float2) - sp.N2) / (2 * (int64_t)kFftRows * sp.N2));
    groups = groups < 1 ? 1 : (groups > 16 ? 16 : groups);
    const size_t smem = ((size_t)sp.N2 + 2 * (size_t)groups * kFftRows * sp.N2) * sizeof(float2);
    {
      static std::mutex mu;
      std::lock_guard<std::mutex> lk(mu);
      AVZ_CUDA_OK(cudaFuncSetAttribute(k_mix_rows_fft<INV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    }
    const int row_blocks = (sp.N1 + groups * kFftRows - 1) / (groups * kFftRows);
    if (used == PP) {
      k_mix_rows_fft<INV><<<dim3(row_blocks, B * PP), kFftThreads, smem, st>>>(Z, W, mul, sp.N1, sp.N2, groups, plan, M);
      AVZ_LAUNCH_OK("k_mix_rows_fft");
      return AVZ_OK;
    }
    for (int q = 0; q < used; ++q) {  // utterance stride PP planes
      k_mix_rows_fft<INV><<<dim3(row_blocks, B), kFftThreads, smem, st>>>(Z + (int64_t)q * M, W, mul, sp.N1, sp.N2, groups, plan,
                                                                  (int64_t)PP * M);
      AVZ_LAUNCH_OK("k_mix_rows_fft");
    }
    return AVZ_OK;
  }
  int Na, Nb;
  sp.row_factors(&Na, &Nb);
  const int64_t M = sp.len();
  const int row_threads = ((sp.N2 + 31) / 32) * 32;
  const int row_blocks = (sp.N1 + kRowsPer - 1) / kRowsPer;
  if (used == PP) {
    k_mix_rows<INV><<<dim3(row_blocks, B * PP), row_threads, sp.row_smem(), st>>>(Z, W, mul, sp.N1, sp.N2, Na, Nb, M);
    AVZ_LAUNCH_OK("k_mix_rows");
    return AVZ_OK;
  }
  for (int q = 0; q < used; ++q) {  // utterance stride PP planes
    k_mix_rows<INV><<<dim3(row_blocks, B), row_threads, sp.row_smem(), st>>>(Z + (int64_t)q * M, W, mul, sp.N1, sp.N2, Na,
                                                                            Nb, (int64_t)PP * M);
    AVZ_LAUNCH_OK("k_mix_rows");
  }
  return AVZ_OK;
}

// Chirp-z plan of one signal length: c[n] = exp(-i pi n^2 / L) (n^2 reduced mod 2L in integers, angle in float64) and
// the spectrum of the wrapped kernel conj(c)[|m|], computed once on the device by the native passes, [k1][k2] order.
struct BluePlan {
  int device;
  int64_t L;
  Split sp;
  float2* chirp;
  float2* bhat;
};
std::mutex g_blue_mu;
std::vector<BluePlan*> g_blue_plans;

int blue_plan_for(int64_t L, const Split& sp, cudaStream_t st, const BluePlan** out) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return set_error(AVZ_ENOGPU, "cudaGetDevice: %s", cudaGetErrorString(e));
  std::lock_guard<std::mutex> lk(g_blue_mu);
  for (const BluePlan* p : g_blue_plans)
    if (p->device == dev && p->L == L) {
      *out = p;
      return AVZ_OK;
    }
  const int64_t M = sp.len();
  const float2* W = nullptr;
  int rc = mix_table_for(M, &W);
  if (rc != AVZ_OK) return rc;
  std::vector<float2> c((size_t)L), kb((size_t)M, make_float2(0.f, 0.f));
  const double pi = 3.141592653589793238462643383279;
  for (int64_t n = 0; n < L; ++n) {
    const int64_t r = (int64_t)(((unsigned long long)n * (unsigned long long)n) % (unsigned long long)(2 * L));
    const double ang = pi * (double)r / (double)L;
    const float cr = (float)cos(ang), ci = (float)sin(ang);
    c[(size_t)n] = make_float2(cr, -ci);
    kb[(size_t)n] = make_float2(cr, ci);
    if (n > 0) kb[(size_t)(M - n)] = make_float2(cr, ci);
  }
  void *dc = nullptr, *dk = nullptr, *dh = nullptr;
  AVZ_CUDA_OK(cudaMalloc(&dc, (size_t)L * sizeof(float2)));
  AVZ_CUDA_OK(cudaMalloc(&dk, (size_t)M * sizeof(float2)));
  AVZ_CUDA_OK(cudaMalloc(&dh, (size_t)M * sizeof(float2)));
  AVZ_CUDA_OK(cudaMemcpy(dc, c.data(), (size_t)L * sizeof(float2), cudaMemcpyHostToDevice));
  AVZ_CUDA_OK(cudaMemcpy(dk, kb.data(), (size_t)M * sizeof(float2), cudaMemcpyHostToDevice));
  rc = set_mix_attrs(sp);
  if (rc != AVZ_OK) return rc;
  const int col_blocks = (sp.N2 + kCols - 1) / kCols;
  k_mix_cols_fwd<kLoadPlane><<<dim3(col_blocks, 1, 1), kColThreads, sp.col_smem(), st>>>(
      nullptr, (const float2*)dk, nullptr, (float2*)dh, W, 1, 1, sp.N1, sp.lg, sp.N2, M, M);
  AVZ_LAUNCH_OK("k_mix_cols_fwd<plane>");
  rc = launch_rows<false>((float2*)dh, W, nullptr, sp, 1, 1, 1, st);
  if (rc != AVZ_OK) return rc;
  AVZ_CUDA_OK(cudaStreamSynchronize(st));   // the plan may be used from any stream afterwards
  AVZ_CUDA_OK(cudaFree(dk));
  BluePlan* p = new BluePlan{dev, L, sp, (float2*)dc, (float2*)dh};
  g_blue_plans.push_back(p);
  *out = p;
  return AVZ_OK;
}

}  // namespace
}  // namespace avz

#include "avz_mixer_cluster.cuh"   // k_mix_cluster: the whole mixer of one utterance inside one 8-CTA cluster

namespace avz {
namespace {
int farfield_mix(const float* src, const double* delays_host, int B, int S, int64_t L, double fs, float peak_eps, float* mix,
                 float* tgt, float* itf, void* ws, void* stream, bool allow_cluster);
}  // namespace
}  // namespace avz

extern "C" {

int64_t avz_farfield_mix_ws_bytes(int B, int S, int64_t L) {
  using namespace avz;
  if (B <= 0 || S < 1 || S > kMaxSrc) return 0;
  const int PP = ((S + 1) / 2 > 2) ? (S + 1) / 2 : 2;
  Split sp;
  if (split_length(L, &sp.N1, &sp.lg, &sp.N2)) return (int64_t)B * PP * L * (int64_t)sizeof(float2) + (int64_t)B * 16;
  if (!bluestein_split(L, &sp)) return 0;
  return (int64_t)B * PP * (sp.len() + L) * (int64_t)sizeof(float2) + (int64_t)B * 16;
}

int avz_farfield_mix_f32(const float* src, const double* delays_host, int B, int S, int64_t L, double fs,
                         float peak_eps, float* mix, float* tgt, float* itf, void* ws, void* stream) {
  return avz::farfield_mix(src, delays_host, B, S, L, fs, peak_eps, mix, tgt, itf, ws, stream, true);
}

int avz_farfield_mix_passes_f32(const float* src, const double* delays_host, int B, int S, int64_t L, double fs,
                                float peak_eps, float* mix, float* tgt, float* itf, void* ws, void* stream) {
  return avz::farfield_mix(src, delays_host, B, S, L, fs, peak_eps, mix, tgt, itf, ws, stream, false);
}

}  // extern "C"

namespace avz {
namespace {
int farfield_mix(const float* src, const double* delays_host, int B, int S, int64_t L, double fs, float peak_eps, float* mix,
                 float* tgt, float* itf, void* ws, void* stream, bool allow_cluster) {
  if (!src || !delays_host || !mix || !tgt || !itf || !ws) return set_error(AVZ_EINVAL, "avz_farfield_mix_f32: null pointer");
  if (S < 1 || S > kMaxSrc) return set_error(AVZ_EINVAL, "avz_farfield_mix_f32: S=%d out of range (1..%d)", S, kMaxSrc);
  if (!(fs > 0.0)) return set_error(AVZ_EINVAL, "avz_farfield_mix_f32: fs must be positive");
  cudaStream_t st = (cudaStream_t)stream;
  const int P = (S + 1) / 2;
  const int PP = P > 2 ? P : 2;
  if (B <= 0 || (int64_t)B * PP > 65535)
    return set_error(AVZ_EINVAL, "avz_farfield_mix_f32: B=%d out of range (B * max(2, ceil(S/2)) <= 65535)", B);
  MixParams prm;
  prm.S = S;
  for (int s = 0; s < kMaxSrc; ++s) {
    prm.c1[s] = s < S ? delays_host[2 * s] * fs / (double)L : 0.0;
    prm.c2[s] = s < S ? delays_host[2 * s + 1] * fs / (double)L : 0.0;
    // symmetric pair (far-field delays +-(d/2) cos(theta) / c; numpy's cos(theta - pi) and -cos(theta) differ in the last
    // bit): the phase difference of treating them as exact negatives is < 1e-11 rad at any bin, far below float32
    prm.sym[s] = (fabs(prm.c2[s] + prm.c1[s]) <= 1e-12 * fabs(prm.c1[s])) ? 1 : 0;
  }
  if (allow_cluster && S <= 4) {   // two complex planes: what one cluster holds on chip
    const ClusterPlan* cp = nullptr;
    const int rc = cluster_plan_for(L, &cp);
    if (rc != AVZ_OK) return rc;
    if (cp != nullptr) return launch_mix_cluster(cp, src, B, S, prm, peak_eps, mix, tgt, itf, st);
  }
  Split sp;
  const bool native = split_length(L, &sp.N1, &sp.lg, &sp.N2);
  const BluePlan* bp = nullptr;
  if (!native) {
    if (!bluestein_split(L, &sp))
      return set_error(AVZ_EINVAL,
                       "avz_farfield_mix_f32: L=%lld unsupported (neither L = 2^a * N2 with 2^a <= %d, N2 <= %d, nor "
                       "2L-1 <= %lld for the chirp-z path)", (long long)L, kMaxN1, kMaxN2, (long long)kMaxN1 * kMaxN2);
    const int rc = blue_plan_for(L, sp, st, &bp);
    if (rc != AVZ_OK) return rc;
  }
  const int64_t M = sp.len();   // transform length (= L on the native path)
  const float2* W = nullptr;
  int rc = mix_table_for(M, &W);
  if (rc != AVZ_OK) return rc;
  float2* Z = (float2*)ws;                                   // [B][PP][M]
  float2* SP = native ? Z : Z + (int64_t)B * PP * M;         // Bluestein: spectra [B][PP][L] in natural bin order
  unsigned* peak = (unsigned*)((char*)ws + (int64_t)B * PP * (native ? L : M + L) * (int64_t)sizeof(float2));
  rc = set_mix_attrs(sp);
  if (rc != AVZ_OK) return rc;
  const size_t col_smem = sp.col_smem();
  const int col_blocks = (sp.N2 + kCols - 1) / kCols;
  AVZ_CUDA_OK(cudaMemsetAsync(peak, 0, (size_t)B * sizeof(unsigned), st));
  if (native) {
    k_mix_cols_fwd<kLoadPair><<<dim3(col_blocks, P, B), kColThreads, col_smem, st>>>(src, nullptr, nullptr, Z, W, S, PP,
                                                                                    sp.N1, sp.lg, sp.N2, L, L);
    AVZ_LAUNCH_OK("k_mix_cols_fwd");
    if ((rc = launch_rows<false>(Z, W, nullptr, sp, B, P, PP, st)) != AVZ_OK) return rc;
    k_mix_combine<<<dim3((unsigned)((L + 255) / 256), B), 256, 0, st>>>(Z, prm, PP, sp.N1, sp.N2, L);
    AVZ_LAUNCH_OK("k_mix_combine");
    if ((rc = launch_rows<true>(Z, W, nullptr, sp, B, 2, PP, st)) != AVZ_OK) return rc;
    k_mix_cols_inv<kStoreReal><<<dim3(col_blocks, 2, B), kColThreads, col_smem, st>>>(Z, W, nullptr, nullptr, mix, tgt, itf,
                                                                                     peak, PP, sp.N1, sp.lg, sp.N2, L, L);
    AVZ_LAUNCH_OK("k_mix_cols_inv");
  } else {
    // Chirp-z (Bluestein): DFT_L(u)[k] = c[k] * sum_n (u[n] c[n]) conj(c)[k - n],  c[n] = exp(-i pi n^2 / L); the
    // convolution runs as a cyclic one of length M >= 2L-1 on the native passes.  Inverse: x = conj(DFT_L(conj Y)) / L.
    k_mix_cols_fwd<kLoadBluePair><<<dim3(col_blocks, P, B), kColThreads, col_smem, st>>>(src, nullptr, bp->chirp, Z, W, S,
                                                                                        PP, sp.N1, sp.lg, sp.N2, M, L);
    AVZ_LAUNCH_OK("k_mix_cols_fwd<chirp>");
    if ((rc = launch_rows<false>(Z, W, nullptr, sp, B, P, PP, st)) != AVZ_OK) return rc;
    if ((rc = launch_rows<true>(Z, W, bp->bhat, sp, B, P, PP, st)) != AVZ_OK) return rc;
    k_mix_cols_inv<kStoreBlueSpec><<<dim3(col_blocks, P, B), kColThreads, col_smem, st>>>(
        Z, W, bp->chirp, SP, nullptr, nullptr, nullptr, nullptr, PP, sp.N1, sp.lg, sp.N2, M, L);
    AVZ_LAUNCH_OK("k_mix_cols_inv<spec>");
    k_mix_combine<<<dim3((unsigned)((L + 255) / 256), B), 256, 0, st>>>(SP, prm, PP, 1, (int)L, L);
    AVZ_LAUNCH_OK("k_mix_combine");
    k_mix_cols_fwd<kLoadBlueSpec><<<dim3(col_blocks, 2, B), kColThreads, col_smem, st>>>(nullptr, SP, bp->chirp, Z, W, S,
                                                                                        PP, sp.N1, sp.lg, sp.N2, M, L);
    AVZ_LAUNCH_OK("k_mix_cols_fwd<spec>");
    if ((rc = launch_rows<false>(Z, W, nullptr, sp, B, 2, PP, st)) != AVZ_OK) return rc;
    if ((rc = launch_rows<true>(Z, W, bp->bhat, sp, B, 2, PP, st)) != AVZ_OK) return rc;
    k_mix_cols_inv<kStoreBlueReal><<<dim3(col_blocks, 2, B), kColThreads, col_smem, st>>>(
        Z, W, bp->chirp, nullptr, mix, tgt, itf, peak, PP, sp.N1, sp.lg, sp.N2, M, L);
    AVZ_LAUNCH_OK("k_mix_cols_inv<real>");
  }
  if (peak_eps >= 0.f) {
    k_mix_scale<<<dim3(64, B), 256, 0, st>>>(mix, tgt, itf, peak, peak_eps, L);
    AVZ_LAUNCH_OK("k_mix_scale");
  }
  return AVZ_OK;
}
}  // namespace
}  // namespace avz

extern "C" {

int avz_pcm16_to_f32(const int16_t* pcm, int64_t n, float* out, void* stream) {
  using namespace avz;
  if (!pcm || !out) return set_error(AVZ_EINVAL, "avz_pcm16_to_f32: null pointer");
  if (n < 0) return set_error(AVZ_EINVAL, "avz_pcm16_to_f32: n < 0");
  if (n == 0) return AVZ_OK;
  if (((uintptr_t)pcm & 15) || ((uintptr_t)out & 15)) return set_error(AVZ_EINVAL, "avz_pcm16_to_f32: buffers must be 16-byte aligned");
  const int64_t want = (n / 8 + 255) / 256 + 1;
  const int grid = (int)(want < (int64_t)num_sms() * 16 ? want : (int64_t)num_sms() * 16);
  k_pcm16_to_f32<<<grid, 256, 0, (cudaStream_t)stream>>>(pcm, n, out);
  AVZ_LAUNCH_OK("k_pcm16_to_f32");
  return AVZ_OK;
}

int avz_f32_to_pcm16(const float* x, int64_t n, int16_t* pcm, void* stream) {
  using namespace avz;
  if (!pcm || !x) return set_error(AVZ_EINVAL, "avz_f32_to_pcm16: null pointer");
  if (n < 0) return set_error(AVZ_EINVAL, "avz_f32_to_pcm16: n < 0");
  if (n == 0) return AVZ_OK;
  if (((uintptr_t)pcm & 15) || ((uintptr_t)x & 15)) return set_error(AVZ_EINVAL, "avz_f32_to_pcm16: buffers must be 16-byte aligned");
  const int64_t want = (n / 8 + 255) / 256 + 1;
  const int grid = (int)(want < (int64_t)num_sms() * 16 ? want : (int64_t)num_sms() * 16);
  k_f32_to_pcm16<<<grid, 256, 0, (cudaStream_t)stream>>>(x, n, pcm);
  AVZ_LAUNCH_OK("k_f32_to_pcm16");
  return AVZ_OK;
}

}  // extern "C"

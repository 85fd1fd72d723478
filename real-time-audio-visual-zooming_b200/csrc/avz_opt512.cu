// n_fft = 512 fast path (the BASELINE shape: 512 / hop 128, also hop 256): register-resident warp FFTs
// (avz_fft512.cuh), one warp per run of consecutive frames, no block-level barrier on the per-frame path.
//
//   k512_ibm        STFT(tgt), STFT(int) packed in one complex transform -> IBM bits (float32 decision; near
//                   ties appended to a list)
//   k512_ibm_fixup  one warp per listed bin: float64 direct DFT, flips the bits float32 got wrong
//   k512_cov        STFT(mic0), STFT(mic1) packed -> mask-weighted 2x2 covariance partial sums in registers
//   k512_apply      STFT(mix) -> beamform + post-filter -> two frames per inverse transform -> window ->
//                   overlap-add in registers -> / sum w^2 -> coalesced stores (+ per-utterance peak)
//
// Each lane keeps a sliding window of raw samples in registers: a new frame costs HOP/32 coalesced 128-byte
// loads per channel, every input sample is fetched once per warp run, and nothing is staged in shared memory
// except the single FFT transposition.  Replaces rt_av_zoom/core/oracle_debug.py:42-94.
#include <cstdio>
#include <cstdlib>

#include "avz_common.cuh"
#include <cooperative_groups.h>

#include "avz_fft512.cuh"

namespace avz {
namespace o512 {

using f512::Lane;
namespace cg = cooperative_groups;

constexpr int kN = 512;
constexpr int kF = 257;
constexpr int kFW = 9;
constexpr int kFP = 288;     // padded bins for partial sums (matches Geo<512>::FP)
constexpr int kWarps = 4;    // warps per CTA
#ifndef AVZ_IBM_FULLTW
#define AVZ_IBM_FULLTW 1     // k512_ibm: unfactored twiddles + per-lane radix-2 constants; needs AVZ_MINB_IBM <= 3 (168 registers)
#endif
#ifndef AVZ_APPLY_FULLTW
#define AVZ_APPLY_FULLTW 0   // k512_apply<KEPT>: unfactored twiddles in the inverse transform (needs AVZ_MINB_APPLY_KEPT 2)
#endif
#ifndef AVZ_MINB_STREAM
#define AVZ_MINB_STREAM 4
#endif
#ifndef AVZ_COV_FULLTW
#define AVZ_COV_FULLTW 1     // k512_cov: 15 unfactored transposition twiddles (+18 registers, -36 instructions per frame)
#endif
#ifndef AVZ_MINB_IBM
#define AVZ_MINB_IBM 3
#endif
#ifndef AVZ_MINB_COV
#define AVZ_MINB_COV 2
#endif
#ifndef AVZ_MINB_APPLY
#define AVZ_MINB_APPLY 2
#endif
#ifndef AVZ_MINB_APPLY_KEPT
#define AVZ_MINB_APPLY_KEPT 3
#endif

// Sliding window of raw samples of two signals (16 rows of 32 lanes = one 512-sample frame each).
// Sample index of row r of frame t for this lane: t * HOP - 256 + 32 r + lane; outside [0, L) it reads as zero
// (scipy boundary='zeros' and padded=True).  L < 2^31 (checked on the host).
template <int HOP>
struct Window2 {
  static constexpr int NR = HOP / 32;   // new rows per frame
  float a[16], b[16];
  float na[NR], nb[NR];                 // rows of the next frame (t + 1), already requested
  float fa[NR], fb[NR];                 // rows of the frame after that (t + 2): loads are kept two frames ahead because
                                        // under load a DRAM round trip outlasts one frame of this warp's compute

  __device__ __forceinline__ void load_rows(float (&ra)[NR], float (&rb)[NR], const float* __restrict__ xa,
                                            const float* __restrict__ xb, int L, int t_new, int lane) {
    const int s0 = t_new * HOP + kN / 2 - HOP;    // first sample of the hop that frame t_new adds
    if (s0 >= 0 && s0 + HOP <= L) {               // warp-uniform: the whole hop is inside the signal
      const float* pa = xa + s0 + lane;
      const float* pb = xb + s0 + lane;
#pragma unroll
      for (int i = 0; i < NR; ++i) {
        ra[i] = __ldg(pa + 32 * i);
        rb[i] = __ldg(pb + 32 * i);
      }
    } else {
#pragma unroll
      for (int i = 0; i < NR; ++i) {
        const int idx = s0 + lane + 32 * i;
        const bool ok = (unsigned)idx < (unsigned)L;
        ra[i] = ok ? __ldg(xa + idx) : 0.f;
        rb[i] = ok ? __ldg(xb + idx) : 0.f;
      }
    }
  }
  // all 16 rows of frame t, and the request for frame t + 1
  __device__ __forceinline__ void load_all(const float* __restrict__ xa, const float* __restrict__ xb, int L, int t,
                                           int lane) {
    const int base = t * HOP - kN / 2 + lane;
#pragma unroll
    for (int r = 0; r < 16; ++r) {
      const int i = base + 32 * r;
      const bool ok = (unsigned)i < (unsigned)L;
      a[r] = ok ? __ldg(xa + i) : 0.f;
      b[r] = ok ? __ldg(xb + i) : 0.f;
    }
    load_rows(na, nb, xa, xb, L, t + 1, lane);
  }
  // while frame t is being processed: request the rows frame t + 2 will add (t_next = t + 1 is already in flight)
  __device__ __forceinline__ void prefetch(const float* __restrict__ xa, const float* __restrict__ xb, int L,
                                           int t_next, int lane) {
    load_rows(fa, fb, xa, xb, L, t_next + 1, lane);
  }
  // move on to frame t + 1
  __device__ __forceinline__ void advance() {
#pragma unroll
    for (int r = 0; r < 16 - NR; ++r) {
      a[r] = a[r + NR];
      b[r] = b[r + NR];
    }
#pragma unroll
    for (int i = 0; i < NR; ++i) {
      a[16 - NR + i] = na[i];
      b[16 - NR + i] = nb[i];
      na[i] = fa[i];
      nb[i] = fb[i];
    }
  }
  // windowed complex frame a + i b
  __device__ __forceinline__ void frame(float2 (&v)[16], const float (&w)[16]) const {
#pragma unroll
    for (int r = 0; r < 16; ++r) v[r] = cscale(make_float2(a[r], b[r]), w[r]);
  }
};

// Window of this lane's 16 rows, times the lane sign the transform wants on its time-domain side.
// The analysis scale 1/sum(w) = 2/N is NOT applied here: it is folded into whatever consumes the spectrum.
__device__ __forceinline__ void load_window(float (&w)[16], const float* __restrict__ win, const Lane& ln) {
#pragma unroll
  for (int r = 0; r < 16; ++r) w[r] = win[32 * r + ln.lane] * ln.sign;
}

// bin of lo[j] for this lane
__device__ __forceinline__ int bin_lo(const Lane& ln, int j) { return ln.k1 + 16 * j + 128 * ln.h; }

// This lane's 8 mask bits (+ Nyquist) of one frame: words 4h .. 4h+3 hold bins 128 h .. 128 h + 127.
struct FrameBits {
  uint32_t w[4], ny;
  __device__ __forceinline__ void load(const uint32_t* __restrict__ bits_t, int h) {
#pragma unroll
    for (int i = 0; i < 4; ++i) w[i] = __ldg(bits_t + 4 * h + i);
    ny = __ldg(bits_t + 8);
  }
  __device__ __forceinline__ bool bit(int j, int k1) const { return (w[j >> 1] >> (k1 + 16 * (j & 1))) & 1u; }
};

// The same bits in (slot j, lane) order: word j of a frame is the warp ballot of slot j, i.e. bit `lane` of word j is the
// noise bit of bin k1 + 16 j + 128 h.  k512_ibm writes this copy next to the k-ordered one when the sparse kept spectrum
// is in use: a slot's keep mask over the lanes (what the compaction needs) is then ~word[j], no ballot or bit shuffling.
struct LaneBits {
  uint32_t w[8], ny;
  __device__ __forceinline__ void load(const uint32_t* __restrict__ lane_bits_t, const uint32_t* __restrict__ bits_t) {
    const uint4 a = __ldg(reinterpret_cast<const uint4*>(lane_bits_t));
    const uint4 b = __ldg(reinterpret_cast<const uint4*>(lane_bits_t) + 1);
    w[0] = a.x, w[1] = a.y, w[2] = a.z, w[3] = a.w, w[4] = b.x, w[5] = b.y, w[6] = b.z, w[7] = b.w;
    ny = __ldg(bits_t + 8);
  }
  __device__ __forceinline__ bool bit(int j, int lane) const { return (w[j] >> lane) & 1u; }
  // lanes of slot j whose bin is kept (noise bit clear); slot (0, lane 0) carries DC and Nyquist: kept if either is
  __device__ __forceinline__ uint32_t keep(int j) const { return ~w[j] | ((j == 0) ? (~ny & 1u) : 0u); }
};

// ------------------------------------------------------------------------------------------
// IBM bits
// ------------------------------------------------------------------------------------------
// Exact decision |S_int[k]| > |S_tgt[k]| of one frame in float64 (direct DFT; the whole warp cooperates).
template <int N>
__device__ __forceinline__ bool ibm_exact512(const float* __restrict__ tgt, const float* __restrict__ itf, int64_t L,
                                             int64_t start, int k, const Tables& tb, int lane) {
  double tr = 0, ti = 0, ir = 0, ii = 0;
#pragma unroll   // N / 32 = 16 independent iterations: all their loads in flight at once (the kernel is latency-bound)
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

// Near-ties are not resolved inline: the kernel records the float32 decision and appends (b, t, k) to a list;
// k512_ibm_fixup then recomputes exactly those bins in float64 (one warp per entry, massively parallel) and
// flips the bits that differ.  If the list overflows, the fix-up kernel rechecks every bin instead.
struct AmbList {
  unsigned long long* entries;  // (b << 32) | (t << 9) | k
  unsigned int* count;
  unsigned int cap;
};

// With Z = FFT(tgt + i int):  |S_int|^2 - |S_tgt|^2 = -Re(Z[k] Z[N-k])  (S_tgt = (Z[k] + conj Z[N-k])/2,
// S_int = -i (Z[k] - conj Z[N-k])/2), so the IBM bit is  Re(Z[k] Z[N-k]) < 0  - no unpacking needed.
// c carries an error of at most delta (|Z[k]| + |Z[N-k]|) + delta^2 when each bin is off by at most delta, hence
// "too close to call" is  c^2 <= delta^2 (4 (|Z[k]|^2 + |Z[N-k]|^2) + 2 delta^2).  Only a frame whose INPUT is
// exactly zero (`live` false) is a true tie everywhere (bit 0, nothing to recheck); a bin that merely comes out
// as 0 + 0i in float32 (cancellation at the rounding floor) is as uncertain as any other small value.
// (k1, k2) = (4 d2, 2 d2^2) for a live frame, (-1, -1) for an all-zero one (the test then never fires).
__device__ __forceinline__ void ibm_decide(float2 z, float2 m, float k1, float k2, bool& bit, bool& amb) {
  const float c = fmaf(z.x, m.x, -z.y * m.y);
  const float s = fmaf(z.x, z.x, fmaf(z.y, z.y, fmaf(m.x, m.x, m.y * m.y)));
  bit = c < 0.f;
  amb = c * c <= fmaf(k1, s, k2);
}

template <int HOP>
__global__ void __launch_bounds__(kWarps * 32, AVZ_MINB_IBM)
k512_ibm(const float* __restrict__ tgt, const float* __restrict__ itf, int L, int T, int frames_per_cta,
         uint32_t* __restrict__ ibm_bits, uint32_t* __restrict__ lane_bits, AmbList amb_list, float tol2, Tables tb) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float2* sm = reinterpret_cast<float2*>(smem_raw) + (size_t)(threadIdx.x >> 5) * f512::kSmemComplex;
  Lane ln;
  ln.init(tb.tw);
#if AVZ_IBM_FULLTW
  ln.init_full(tb.tw);
#endif
  const int lane = ln.lane, warp = warp_id_uniform();
  const int b = blockIdx.y;
  const float* tg = tgt + (int64_t)b * L;
  const float* it = itf + (int64_t)b * L;
  float w[16];
  load_window(w, tb.win, ln);

  const int c0 = blockIdx.x * frames_per_cta, c1 = min(T, c0 + frames_per_cta);
  const int per = (c1 - c0 + kWarps - 1) / kWarps;
  const int ta = c0 + warp * per, tb_ = min(c1, ta + per);
  if (ta >= tb_) return;

  Window2<HOP> win;
  win.load_all(tg, it, L, ta, lane);
#pragma unroll 1
  for (int t = ta; t < tb_; ++t) {
    if (t + 1 < tb_) win.prefetch(tg, it, L, t + 1, lane);
    float2 v[16];
    win.frame(v, w);
    u64 e2p = pk2(0.f, 0.f);
#pragma unroll
    for (int r = 0; r < 16; ++r) e2p = fma2(pk2(v[r]), pk2(v[r]), e2p);
    float e2 = warp_sum(up2(e2p).x + up2(e2p).y);
#if AVZ_IBM_FULLTW
    f512::forward_full(v, sm, ln);
#else
    f512::forward(v, sm, ln);
#endif
    float2 mir[8];
    f512::mirror_of_low(v, mir, ln);
    // float32 FFT error bound of one bin: delta = tol * sqrt(e2)  (e2 = sum |frame|^2, so 512 e2 = sum |Z|^2
    // and sqrt(e2) is the rms bin magnitude; the spectrum here is unscaled).
    const float d2 = tol2 * e2;
    const bool live = e2 > 0.f;   // any non-zero input sample in either reference
    const float k1 = live ? 4.f * d2 : -1.f, k2 = live ? 2.f * d2 * d2 : -1.f;
    unsigned ballots[8];
    unsigned my_amb = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      bool bit, amb;
      ibm_decide(v[j], mir[j], k1, k2, bit, amb);
      ballots[j] = __ballot_sync(kFull, bit);
      my_amb |= amb ? (1u << j) : 0u;
    }
    // Nyquist bin 256 = hi[0] of lane 0, its own mirror: same formula with m = z
    bool ny_bit, ny_a;
    ibm_decide(v[8], v[8], k1, k2, ny_bit, ny_a);
    const unsigned ny = __ballot_sync(kFull, ny_bit) & 1u;
    if (lane == 0 && ny_a) my_amb |= 1u << 8;
    if (__any_sync(kFull, my_amb != 0u)) {  // warp-uniform, taken for a minority of frames
      const int cnt = __popc(my_amb);
      int incl = cnt;                        // inclusive prefix sum over lanes
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int up = __shfl_up_sync(kFull, incl, o);
        if (lane >= o) incl += up;
      }
      const int tot = __shfl_sync(kFull, incl, 31);
      unsigned base = 0;
      if (lane == 31) base = atomicAdd(amb_list.count, (unsigned)tot);
      base = __shfl_sync(kFull, base, 31) + (unsigned)(incl - cnt);
      const unsigned long long hdr = ((unsigned long long)b << 32) | ((unsigned long long)t << 9);
      unsigned m = my_amb;
      while (m) {
        const int j = __ffs(m) - 1;
        m &= m - 1;
        if (base < amb_list.cap) amb_list.entries[base] = hdr | (unsigned)(j == 8 ? 256 : bin_lo(ln, j));
        ++base;
      }
    }
    if (lane == 0) {
      uint32_t* o = ibm_bits + ((int64_t)b * T + t) * kFW;
#pragma unroll
      for (int wd = 0; wd < 4; ++wd) {
        o[wd] = __byte_perm(ballots[2 * wd], ballots[2 * wd + 1], 0x5410);       // low halves:  (x & 0xffff) | (y << 16)
        o[4 + wd] = __byte_perm(ballots[2 * wd], ballots[2 * wd + 1], 0x7632);   // high halves: (x >> 16) | (y & 0xffff0000)
      }
      o[8] = ny;
      if (lane_bits != nullptr) {   // the ballots themselves: (slot, lane) order for the sparse kept spectrum
        uint4* lb = reinterpret_cast<uint4*>(lane_bits + ((int64_t)b * T + t) * 8);
        lb[0] = make_uint4(ballots[0], ballots[1], ballots[2], ballots[3]);
        lb[1] = make_uint4(ballots[4], ballots[5], ballots[6], ballots[7]);
      }
    }
    if (t + 1 < tb_) win.advance();
  }
}

// Exact float64 decisions for the listed bins; a stored bit is flipped where float32 got it wrong.
// Each warp takes blocks of kFixBlock consecutive list entries.  Entries of one frame are contiguous in the list (they
// were appended by one warp in one reservation), so the frame's windowed samples are converted to float64 once, kept
// in registers (lane holds samples n = lane + 32 r), and reused for every listed bin of that frame: per bin only the
// 16 twiddle loads, 64 DFMA and the warp reduction remain.
constexpr int kFixBlock = 2;   // short blocks + many warps: the kernel is a latency chain per entry, so spread it wide

__global__ void __launch_bounds__(256)
k512_ibm_fixup(const float* __restrict__ tgt, const float* __restrict__ itf, int64_t L, int T, int hop, int B,
               uint32_t* __restrict__ ibm_bits, uint32_t* __restrict__ lane_bits, AmbList amb_list, Tables tb) {
  const int lane = threadIdx.x & 31;
  const unsigned long long gw = (unsigned long long)blockIdx.x * (blockDim.x >> 5) + (unsigned)warp_id_uniform();
  const unsigned long long nw = (unsigned long long)gridDim.x * (blockDim.x >> 5);
  const unsigned int n = *amb_list.count;
  const bool overflow = n > amb_list.cap;
  const unsigned long long total = overflow ? (unsigned long long)B * T * kF : n;
  float xt[16], xi[16];   // raw samples of the cached frame (float: half the registers of the windowed doubles, so
                          // twice the warps per SM for a kernel that is one latency chain per entry)
  int cur_b = -1, cur_t = -1;
  for (unsigned long long i0 = gw * kFixBlock; i0 < total; i0 += nw * kFixBlock) {
    const unsigned long long i1 = (i0 + kFixBlock < total) ? i0 + kFixBlock : total;
    for (unsigned long long i = i0; i < i1; ++i) {
      int b, t, k;
      if (!overflow) {
        const unsigned long long e = amb_list.entries[i];
        b = (int)(e >> 32);
        t = (int)((e >> 9) & 0x7fffffu);
        k = (int)(e & 511u);
      } else {
        k = (int)(i % kF);
        t = (int)((i / kF) % T);
        b = (int)(i / ((unsigned long long)kF * T));
      }
      if (b != cur_b || t != cur_t) {   // warp-uniform
        cur_b = b;
        cur_t = t;
        const float* tg = tgt + (int64_t)b * L;
        const float* it = itf + (int64_t)b * L;
        const int64_t start = (int64_t)t * hop - kN / 2;
#pragma unroll
        for (int r = 0; r < 16; ++r) {
          const int nn = lane + 32 * r;
          const int64_t idx = start + nn;
          const bool ok = idx >= 0 && idx < L;
          xt[r] = ok ? __ldg(tg + idx) : 0.f;
          xi[r] = ok ? __ldg(it + idx) : 0.f;
        }
      }
      // exp(-2 pi i (lane + 32 r) k / N) = base * step^r: one gathered and one broadcast table load per entry instead
      // of 16 gathers (random 16-byte gathers cost up to 32 L1 wavefronts each and were this kernel's bottleneck);
      // the recurrence adds ~16 ulp of float64 rounding, far inside the decision margin.
      double2 e = tb.tw_d[(lane * k) & (kN - 1)];
      const double2 stp = tb.tw_d[(32 * k) & (kN - 1)];
      double tr = 0, ti = 0, ir = 0, ii = 0;
#pragma unroll
      for (int r = 0; r < 16; ++r) {
        const double w = tb.win_d[lane + 32 * r];
        const double at = w * (double)xt[r], ai = w * (double)xi[r];
        tr = fma(at, e.x, tr);
        ti = fma(at, e.y, ti);
        ir = fma(ai, e.x, ir);
        ii = fma(ai, e.y, ii);
        const double nx = fma(e.x, stp.x, -e.y * stp.y), ny = fma(e.x, stp.y, e.y * stp.x);
        e = make_double2(nx, ny);
      }
      tr = warp_sum(tr);
      ti = warp_sum(ti);
      ir = warp_sum(ir);
      ii = warp_sum(ii);
      const bool exact = (ir * ir + ii * ii) > (tr * tr + ti * ti);
      if (lane == 0) {
        uint32_t* wp = ibm_bits + ((int64_t)b * T + t) * kFW + (k >> 5);
        const bool cur = ((*wp) >> (k & 31)) & 1u;
        if (cur != exact) {
          atomicXor(wp, 1u << (k & 31));
          if (lane_bits != nullptr && k < 256)   // bin k = k1 + 16 j + 128 h sits at bit k1 + 16 h of word j
            atomicXor(lane_bits + ((int64_t)b * T + t) * 8 + ((k & 127) >> 4), 1u << ((k & 15) + 16 * (k >> 7)));
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// masked covariance partial sums
// ------------------------------------------------------------------------------------------
enum { W_BITS = 0, W_MASK = 1 };

// Where mask[b][k][t] lives: the reference layout (B, F, T) has T contiguous, which makes the per-frame reads of a
// warp (one bin per lane) 32 separate sectors per load.  When the kept-spectrum buffer is available its tail holds a
// transposed copy (B, T, kMaskPitch) made by k_mask_transpose, and a frame's 257 weights are 9 contiguous runs.
struct MaskLayout {
  int64_t sb;   // floats between utterances
  int sf;       // floats between bins
  int st;       // floats between frames
  const uint32_t* hdr;   // pass B on a transposed copy it did not make itself: header to verify (else nullptr)
  uint32_t B, T;
};
constexpr int kMaskPitch = 264;   // 257 bins padded to a multiple of 8 floats (32-byte rows)
// The transposed mask copy behind the kept spectrum is announced by a header {magic, B, T, 0} that k_mask_transpose
// writes and the IBM pass A clears: pass B called with mask == NULL ("use the copy pass A staged") verifies it on the
// device and poisons its output with NaN when the copy is not there - a stale or foreign buffer must not be read as a
// mask silently (no host round trip: the call stays asynchronous and graph-capturable).
constexpr uint32_t kMaskMagic = 0x4d41534bu;   // "MASK"
constexpr size_t kMaskHdrBytes = 256;

__global__ void __launch_bounds__(256)
k_mask_transpose(const float* __restrict__ mask, float* __restrict__ mask_t, int T, uint32_t* __restrict__ hdr) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z, k0 = blockIdx.y * 32, t0 = blockIdx.x * 32;
  if (hdr != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && threadIdx.x == 0) {
    hdr[0] = kMaskMagic;
    hdr[1] = gridDim.z;
    hdr[2] = (uint32_t)T;
    hdr[3] = 0u;
  }
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
  const float* src = mask + (int64_t)b * kF * T;
  float* dst = mask_t + (int64_t)b * T * kMaskPitch;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int k = k0 + ty + 8 * i, t = t0 + tx;
    tile[ty + 8 * i][tx] = (k < kF && t < T) ? __ldg(src + (int64_t)k * T + t) : 0.f;
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int t = t0 + ty + 8 * i, k = k0 + tx;
    if (t < T && k < kMaskPitch) dst[(int64_t)t * kMaskPitch + k] = tile[tx][ty + 8 * i];
  }
}

// `spec` (optional): the packed two-mic spectrum of every frame is kept for pass B, so that pass B does not repeat
// the forward transform: [B][T][8][32] float4 = (lo[i].x, lo[i].y, mir[i].x, mir[i].y) of lane l (lane 0's i = 0
// slot carries (DC, Nyquist): its mirror is itself).  Each store instruction writes 512 contiguous bytes.
// This trades 64 B/sample of spare HBM bandwidth for ~30 % fewer instructions on an issue-bound path.
// L2KEEP: the kept spectrum is written with ordinary stores (it is read back out of L2 by the same launch: k512_fused)
// instead of streaming (evict-first) stores.  (b, chunk, chunks) = (blockIdx.y, blockIdx.x, gridDim.x) of the plain launch.
// KMODE (how the spectrum is kept, W_BITS only): 0 every slot; 1 sparse (compacted, see below); 2 dense layout, but a
// 32-byte sector (the slots of two adjacent lanes) whose two bins are both noise-dominated is not written - pass B, which
// must then apply the post-filter 1 - noise mask with the same bits, multiplies whatever sits there by zero: the buffer has
// to hold finite values from the start (zero it once; a NaN left in it would surface in the output, loudly).
enum { KEEP_ALL = 0, KEEP_SPARSE = 1, KEEP_SKIP = 2 };
template <int HOP, int WMODE, bool L2KEEP, int KMODE>
__device__ __forceinline__ void cov_body(unsigned char* smem_raw, int b, int chunk, int chunks,
                                         const float* __restrict__ mix, const uint32_t* __restrict__ ibm_bits,
                                         const float* __restrict__ mask, MaskLayout ml, int L, int T, int frames_per_cta,
                                         float sqrt_eps, float* __restrict__ part, float4* __restrict__ spec,
                                         int64_t spec_utt, const uint32_t* __restrict__ lane_bits, const Tables& tb) {
  float2* sm_all = reinterpret_cast<float2*>(smem_raw);
  float2* sm = sm_all + (size_t)(threadIdx.x >> 5) * f512::kSmemComplex;
  Lane ln;
  ln.init(tb.tw);
#if AVZ_COV_FULLTW
  ln.init_full(tb.tw);   // this kernel runs 2 CTAs/SM on its accumulators anyway: spend spare registers on twiddles
#endif
  const int lane = ln.lane, warp = warp_id_uniform();
  const float* m0 = mix + (int64_t)b * 2 * L;
  const float* m1 = m0 + L;
  float w[16];
  load_window(w, tb.win, ln);

  // accumulators of this lane's 8 low bins (+ Nyquist on lane 0): R00, R11, Re R01, Im R01, sum m.
  // The spectra here are unscaled (x N/2) and un-halved (Y0' = 2 Y0): products are scaled by 1/N^2 at the end.
  float a00[8], a11[8], are[8], aim[8], am_[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) a00[j] = a11[j] = are[j] = aim[j] = am_[j] = 0.f;
  float n00 = 0.f, n11 = 0.f, nre = 0.f, nm = 0.f;

  const int c0 = chunk * frames_per_cta, c1 = min(T, c0 + frames_per_cta);
  const int per = (c1 - c0 + kWarps - 1) / kWarps;
  const int ta = c0 + warp * per, tb_ = min(c1, ta + per);

  if (ta < tb_) {
    Window2<HOP> win;
    win.load_all(m0, m1, L, ta, lane);
    // noise weights of this lane's bins, fetched one frame ahead
    FrameBits nb;
    unsigned pairkeep = 0xffu;
    LaneBits lq;      // (KMODE == KEEP_SPARSE): the same bits in (slot, lane) order
    float nmask[9];
    const uint32_t* bw = (WMODE == W_BITS) ? ibm_bits + ((int64_t)b * T + ta) * kFW : nullptr;
    const uint32_t* lw = (KMODE == KEEP_SPARSE) ? lane_bits + ((int64_t)b * T + ta) * 8 : nullptr;
    const float* mk = (WMODE == W_MASK) ? mask + (int64_t)b * ml.sb + (int64_t)ta * ml.st : nullptr;
    if ((KMODE == KEEP_SPARSE)) {
      lq.load(lw, bw);
    } else if (WMODE == W_BITS) {
      nb.load(bw, ln.h);
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) nmask[j] = __ldg(mk + (int64_t)bin_lo(ln, j) * ml.sf);
      nmask[8] = __ldg(mk + (int64_t)256 * ml.sf);
    }
#pragma unroll 1
    for (int t = ta; t < tb_; ++t) {
      float mw[8], mny;
      if ((KMODE == KEEP_SPARSE)) {
#pragma unroll
        for (int j = 0; j < 8; ++j) mw[j] = lq.bit(j, lane) ? 1.f : 0.f;
        mny = (lq.ny & 1u) ? 1.f : 0.f;
      } else if (WMODE == W_BITS) {
#pragma unroll
        for (int j = 0; j < 8; ++j) mw[j] = nb.bit(j, ln.k1) ? 1.f : 0.f;
        mny = (nb.ny & 1u) ? 1.f : 0.f;
        if (KMODE == KEEP_SKIP) {   // bit j: the sector of lanes (k1 & ~1, k1 | 1) of slot j has a bin pass B will use
          pairkeep = 0u;
#pragma unroll
          for (int j = 0; j < 8; ++j)
            pairkeep |= ((((nb.w[j >> 1] >> ((ln.k1 & ~1) + 16 * (j & 1))) & 3u) != 3u) ? 1u : 0u) << j;
          if (lane < 2 && !(nb.ny & 1u)) pairkeep |= 1u;   // slot (0, lane 0) also carries the Nyquist bin
        }
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) mw[j] = 1.f - nmask[j];
        mny = 1.f - nmask[8];
      }
      if (t + 1 < tb_) {
        win.prefetch(m0, m1, L, t + 1, lane);
        if ((KMODE == KEEP_SPARSE)) {
          // the next frame's bits are fetched after this frame's keep masks have been used (below)
        } else if (WMODE == W_BITS) {
          bw += kFW;
          nb.load(bw, ln.h);
        } else {
          mk += ml.st;
#pragma unroll
          for (int j = 0; j < 8; ++j) nmask[j] = __ldg(mk + (int64_t)bin_lo(ln, j) * ml.sf);
          nmask[8] = __ldg(mk + (int64_t)256 * ml.sf);
        }
      }
      float2 v[16];
      win.frame(v, w);
#if defined(AVZ_COV_NOFFT)   // timing experiment only (profiles/README.md): everything but the transform
#elif AVZ_COV_FULLTW
      f512::forward_full(v, sm, ln);
#else
      f512::forward(v, sm, ln);
#endif
      float2 mir[8];
      f512::mirror_of_low(v, mir, ln);
      if (spec != nullptr) {
        // `sparse` (oracle path, post-filter 1 - noise mask): pass B multiplies every bin whose noise bit is set by zero,
        // so only the bins with a clear bit are kept - compacted to the front of the frame's 4096-byte block in
        // (slot j, lane) order, which both passes derive from the same IBM bits.  About two thirds of the kept-spectrum
        // traffic, which is what bounds this kernel and pass B, disappears.
        float4* sp = spec + (spec_utt * T + t) * 256;
        const unsigned lt = (1u << lane) - 1u;
        int base = 0;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float4 q = make_float4(v[j].x, v[j].y, mir[j].x, mir[j].y);
          if (j == 0 && lane == 0) {  // (DC, Nyquist)
            q.z = v[8].x;
            q.w = v[8].y;
          }
          const unsigned bal = (KMODE == KEEP_SPARSE) ? lq.keep(j) : kFull;    // lanes of this slot that are kept
          const bool keep = (KMODE == KEEP_SKIP) ? ((pairkeep >> j) & 1u) : ((bal >> lane) & 1u);
          float4* dst = (KMODE == KEEP_SPARSE) ? sp + base + __popc(bal & lt) : sp + 32 * j + lane;
          base += __popc(bal);
          if (keep) {
            if (L2KEEP) *dst = q;
            else __stcs(dst, q);     // streaming store: read once by pass B, no reuse before that
          }
        }
      }
      if ((KMODE == KEEP_SPARSE) && t + 1 < tb_) {
        bw += kFW;
        lw += 8;
        lq.load(lw, bw);
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        // Y0' = Z[k] + conj Z[N-k] = 2 Y0 ;  Y1' = -i (Z[k] - conj Z[N-k]) = 2 Y1
        const float2 y0 = caddc(v[j], mir[j]);
        const float2 y1 = up2(add2(pk2(v[j].y, -v[j].x), pk2(mir[j].y, mir[j].x)));
        const float m = mw[j];
        const float ms = (WMODE == W_MASK) ? m + sqrt_eps : m;
        const float2 tt = cscale(y0, ms), uu = cscale(y1, ms);
        const float t0 = tt.x, t1 = tt.y;
        a00[j] = fmaf(t0, y0.x, fmaf(t1, y0.y, a00[j]));
        are[j] = fmaf(t0, y1.x, fmaf(t1, y1.y, are[j]));   // Re(y0 conj y1)
        aim[j] = fmaf(t1, y1.x, fmaf(-t0, y1.y, aim[j]));  // Im(y0 conj y1)
        a11[j] = fmaf(uu.x, y1.x, fmaf(uu.y, y1.y, a11[j]));
        am_[j] += m;
      }
      {  // Nyquist (meaningful on lane 0 only): Y0 = Re hi[0], Y1 = Im hi[0] (not doubled)
        const float ms = (WMODE == W_MASK) ? mny + sqrt_eps : mny;
        n00 = fmaf(ms * v[8].x, v[8].x, n00);
        n11 = fmaf(ms * v[8].y, v[8].y, n11);
        nre = fmaf(ms * v[8].x, v[8].y, nre);
        nm += mny;
      }
      if (t + 1 < tb_) win.advance();
    }
  }
  // per-warp sums -> shared -> fixed-order CTA reduction -> partial buffer [B][chunks][5][kFP]
  const float sc = 1.0f / ((float)kN * (float)kN);      // (2/N)^2 analysis scale x 1/4 for the un-halved products
  const float sc_ny = 4.0f / ((float)kN * (float)kN);   // Nyquist values were not doubled
  float* s_acc = reinterpret_cast<float*>(sm_all + (size_t)kWarps * f512::kSmemComplex);  // [kWarps][5][kFP]
  float* mine = s_acc + (size_t)warp * 5 * kFP;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int k = bin_lo(ln, j);
    mine[0 * kFP + k] = sc * a00[j];
    mine[1 * kFP + k] = sc * a11[j];
    mine[2 * kFP + k] = sc * are[j];
    mine[3 * kFP + k] = sc * aim[j];
    mine[4 * kFP + k] = am_[j];
  }
  if (lane == 0) {
    mine[0 * kFP + 256] = sc_ny * n00;
    mine[1 * kFP + 256] = sc_ny * n11;
    mine[2 * kFP + 256] = sc_ny * nre;
    mine[3 * kFP + 256] = 0.f;
    mine[4 * kFP + 256] = nm;
  }
  __syncthreads();
  float* dst = part + ((int64_t)b * chunks + chunk) * 5 * kFP;
  for (int i = threadIdx.x; i < 5 * kFP; i += kWarps * 32) {
    const int k = i % kFP;
    float s = 0.f;
    if (k < kF) {
#pragma unroll
      for (int ww = 0; ww < kWarps; ++ww) s += s_acc[(size_t)ww * 5 * kFP + i];
    }
    dst[i] = s;
  }
}

// The last pass-A block of utterance u: float64 sum of the chunk partials in k_cov_finalize's order (chunks c = q mod 4
// summed per slice q, slices combined in order) and the closed-form weights of k_mvdr_weights - same arithmetic, same bits.
__device__ __forceinline__ void finalize_weights_utt(const float* __restrict__ part, int u, int chunks, float norm_eps,
                                                     const float2* __restrict__ dvec, const AvzMvdrCfg& cfg,
                                                     float4* __restrict__ R, float* __restrict__ msum, float2* __restrict__ w) {
  for (int k = threadIdx.x; k < kF; k += kWarps * 32) {
    double sl[4][5];
#pragma unroll
    for (int q = 0; q < 4; ++q)
#pragma unroll
      for (int j = 0; j < 5; ++j) sl[q][j] = 0.0;
    for (int c0 = 0; c0 < chunks; c0 += 4)
#pragma unroll
      for (int q = 0; q < 4; ++q)
        if (c0 + q < chunks) {
          const float* p = part + ((int64_t)u * chunks + c0 + q) * 5 * kFP;
#pragma unroll
          for (int j = 0; j < 5; ++j) sl[q][j] += (double)__ldcg(p + j * kFP + k);
        }
    double sj[5];
#pragma unroll
    for (int j = 0; j < 5; ++j) sj[j] = ((sl[0][j] + sl[1][j]) + sl[2][j]) + sl[3][j];
    const double inv = 1.0 / (sj[4] + (double)norm_eps);
    const float4 r = make_float4((float)(sj[0] * inv), (float)(sj[1] * inv), (float)(sj[2] * inv), (float)(sj[3] * inv));
    const int64_t idx = (int64_t)u * kF + k;
    R[idx] = r;
    msum[idx] = (float)sj[4];
    float2 w0, w1;
    mvdr_weights_bin(r, dvec[2 * k], dvec[2 * k + 1], k, cfg, w0, w1);
    w[2 * idx] = w0;
    w[2 * idx + 1] = w1;
  }
}

// Pass A with the per-utterance tail folded in: the block that completes an utterance's last chunk (a counter per
// utterance) sums the partials and solves the 257 2x2 systems - two launches fewer per step, which is what a single
// utterance (BASELINE config 1: a chain of dependent launches of a few microseconds each) is made of.
struct CovTail {
  int* done;               // [B] zeroed by the launcher
  const float2* dvec;
  float4* R;
  float* msum;
  float2* w;
  AvzMvdrCfg cfg;
  float norm_eps;
  uint32_t* shdr;          // header of the kept-spectrum buffer: word 3 records whether the spectrum is sparse
  int sparse;
  const uint32_t* lane_bits;
};
template <int HOP, int KMODE>
__global__ void __launch_bounds__(kWarps * 32, AVZ_MINB_COV)
k512_cov_w(const float* __restrict__ mix, const uint32_t* __restrict__ ibm_bits, int L, int T, int frames_per_cta,
           float* __restrict__ part, float4* __restrict__ spec, CovTail tail, Tables tb) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ int s_last;
  if (tail.shdr != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) tail.shdr[3] = (uint32_t)KMODE;
  cov_body<HOP, W_BITS, false, KMODE>(smem_raw, blockIdx.y, blockIdx.x, gridDim.x, mix, ibm_bits, nullptr,
                                       MaskLayout{0, 0, 0, nullptr, 0u, 0u}, L, T, frames_per_cta, 0.f, part, spec, blockIdx.y,
                                       tail.lane_bits, tb);
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = (atomicAdd(tail.done + blockIdx.y, 1) == (int)gridDim.x - 1);
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  finalize_weights_utt(part, blockIdx.y, gridDim.x, tail.norm_eps, tail.dvec, tail.cfg, tail.R, tail.msum, tail.w);
}

template <int HOP, int WMODE, int KMODE>
__global__ void __launch_bounds__(kWarps * 32, AVZ_MINB_COV)
k512_cov(const float* __restrict__ mix, const uint32_t* __restrict__ ibm_bits, const float* __restrict__ mask,
         MaskLayout ml, int L, int T, int frames_per_cta, float sqrt_eps, float* __restrict__ part,
         float4* __restrict__ spec, uint32_t* __restrict__ shdr, const uint32_t* __restrict__ lane_bits, Tables tb) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  if (WMODE == W_BITS && shdr != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) shdr[3] = (uint32_t)KMODE;
  cov_body<HOP, WMODE, false, KMODE>(smem_raw, blockIdx.y, blockIdx.x, gridDim.x, mix, ibm_bits, mask, ml, L, T, frames_per_cta,
                                      sqrt_eps, part, spec, blockIdx.y, lane_bits, tb);
}

// ------------------------------------------------------------------------------------------
// beamform + post-filter + inverse + overlap-add
// ------------------------------------------------------------------------------------------
enum { GAIN_NONE = 0, GAIN_BITS = 1, GAIN_FLOOR = 2, GAIN_MASK = 3 };

#ifndef AVZ_STAGES
#define AVZ_STAGES 2
#endif
constexpr int kStages = AVZ_STAGES;   // frames of kept spectrum in flight per warp (TMA ring)
template <int HOP>
__host__ __device__ constexpr size_t apply_fixed_smem() {
  return (size_t)kWarps * f512::kSmemComplex * sizeof(float2) + 260 * sizeof(float4) +
         2 * (size_t)kWarps * (16 - HOP / 32) * 32 * sizeof(float);
}
template <int HOP, bool KEPT>
__host__ __device__ constexpr size_t apply_smem_bytes() {
  return KEPT ? (apply_fixed_smem<HOP>() + 127) / 128 * 128 + (size_t)kWarps * kStages * (4096 + 8)
              : apply_fixed_smem<HOP>();
}

// KEPT: the packed mix spectrum of every frame was stored by k512_cov (`spec`); no forward transform here: frames
// are staged into a per-warp shared-memory ring by TMA bulk copies (cp.async.bulk + mbarrier), kStages deep.
// (b, bx) = (blockIdx.y, blockIdx.x) of the plain launch; spec_utt: utterance slot of the kept spectrum (= b in the
// plain launch, a ring slot in k512_fused).  The mbarriers of the TMA ring are initialised once per CTA (`init_bar`) and
// keep counting phases across the tasks of a persistent kernel: bit `stage` of `phase_bits` is the parity the next wait
// on that stage must use.
template <int HOP, bool KEPT, bool SPARSE>
__device__ __forceinline__ void apply_body(unsigned char* smem_raw, int b, int bx, bool init_bar, uint32_t& phase_bits,
                                           int64_t spec_utt,
                                           const float* __restrict__ mix, const float4* __restrict__ spec,
                                           const float2* __restrict__ wgt, const uint32_t* __restrict__ ibm_bits,
                                           const float* __restrict__ mask, MaskLayout ml, int gain_mode, float post_floor,
                                           int L, int T, int blocks_per_cta, float* __restrict__ out,
                                           float* __restrict__ peak, int cluster_norm, float peak_eps,
                                           const uint32_t* __restrict__ shdr, const uint32_t* __restrict__ lane_bits,
                                           const Tables& tb) {
  constexpr int R = kN / HOP;        // frames overlapping one hop-block
  constexpr int NR = HOP / 32;       // rows per hop-block
  constexpr int TAIL = 16 - NR;      // rows still open after a frame's first block is emitted
  float2* sm_all = reinterpret_cast<float2*>(smem_raw);
  float2* sm = sm_all + (size_t)(threadIdx.x >> 5) * f512::kSmemComplex;
  float4* s_ab = reinterpret_cast<float4*>(sm_all + (size_t)kWarps * f512::kSmemComplex);  // [kF] (a.x,a.y,b.x,b.y)
  float* s_head = reinterpret_cast<float*>(s_ab + 260);         // [kWarps][TAIL][32] first open blocks of a run
  float* s_tail = s_head + kWarps * TAIL * 32;                  // [kWarps][TAIL][32] blocks left open at its end
  // KEPT: per-warp ring of kStages frames of kept spectrum (4096 B each), filled by TMA bulk copies
  constexpr size_t kRingOff = (apply_fixed_smem<HOP>() + 127) / 128 * 128;
  float4* s_ring = reinterpret_cast<float4*>(smem_raw + kRingOff) + (size_t)(threadIdx.x >> 5) * kStages * 256;
  uint64_t* s_bar = reinterpret_cast<uint64_t*>(smem_raw + kRingOff + (size_t)kWarps * kStages * 4096) +
                    (size_t)(threadIdx.x >> 5) * kStages;
  __shared__ float s_peak[kWarps];

  Lane ln;
  ln.init(tb.tw);
#if AVZ_COV_FULLTW
  if (!KEPT || AVZ_APPLY_FULLTW) ln.init_full(tb.tw);   // !KEPT: same twiddles as k512_cov, so recomputed and kept spectra are bit-identical
#endif
  const int lane = ln.lane, warp = warp_id_uniform();
  const float* m0 = mix + (int64_t)b * 2 * L;
  const float* m1 = m0 + L;
  // S[k] = conj(w0) Y0 + conj(w1) Y1 = a[k] Z[k] + b[k] conj(Z[N-k]),  a = (conj w0 - i conj w1)/2,
  // b = (conj w0 + i conj w1)/2.  The transforms here are unscaled: the analysis scale 2/N and the synthesis
  // factor (irfft's 1/N times sum(w) = N/2, i.e. 1/2) are folded into a and b: overall 1/N.
  {
    // a staged mask copy that is not there (header mismatch) poisons the weights: the output is NaN, not garbage
    // ... and so does a kept spectrum whose layout (dense / sparse) is not the one this call was told to read
    const bool staged_ok = (ml.hdr == nullptr || (ml.hdr[0] == kMaskMagic && ml.hdr[1] == ml.B && ml.hdr[2] == ml.T)) &&
                           (!KEPT || shdr == nullptr ||
                            (SPARSE ? shdr[3] == 1u : (shdr[3] == 0u || (gain_mode == GAIN_BITS && shdr[3] == 2u))));
    const float sc = staged_ok ? 0.5f / (float)kN : __int_as_float(0x7fc00000);
    for (int k = threadIdx.x; k < kF; k += kWarps * 32) {
      const float2 w0 = wgt[((int64_t)b * kF + k) * 2 + 0];
      const float2 w1 = wgt[((int64_t)b * kF + k) * 2 + 1];
      s_ab[k] = make_float4(sc * (w0.x - w1.y), sc * (-w0.y - w1.x), sc * (w0.x + w1.y), sc * (-w0.y + w1.x));
    }
  }
  if (KEPT && init_bar && lane == 0) {
#pragma unroll
    for (int st = 0; st < kStages; ++st) mbar_init(s_bar + st, 1);
    mbar_fence_init();
  }
  __syncthreads();

  float hw[16];   // Hann x lane sign: analysis and synthesis window (sign^2 = 1 in the normaliser)
  load_window(hw, tb.win, ln);
  // 1 / sum over the R overlapping frames of w^2, for a block with all R frames present
  float inv_full[NR];
#pragma unroll
  for (int r = 0; r < NR; ++r) {
    float s = 0.f;
#pragma unroll
    for (int q = 0; q < R; ++q) s = fmaf(hw[r + NR * q], hw[r + NR * q], s);
    inv_full[r] = 1.0f / s;
  }

  // Hop-blocks in extended coordinates: block g covers [g HOP, (g+1) HOP); the output is blocks
  // [R/2, R/2 + T - 1); block g sums frames g-R+1 .. g (those that exist).  This CTA owns [G0, G1), split
  // into contiguous runs per warp.  Warp 0 recomputes the R-1 frames before G0 (warm-up) and gets a shorter
  // run to balance that; later warps start cold and their first R-1 blocks are completed at the end from the
  // previous warp's open tail.  Runs are at least R-1 blocks long so a head block only needs one tail.
  const int g_lo = R / 2, g_hi = R / 2 + T - 1;
  const int G0 = g_lo + bx * blocks_per_cta, G1 = min(g_hi, G0 + blocks_per_cta);
  const int nblk = G1 - G0;
  const int per = max(R - 1, (nblk + (R - 1) + kWarps - 1) / kWarps);   // frames per warp incl. warp 0's warm-up
  const int first = max(min(nblk, R - 1), min(nblk, per - (R - 1)));    // warp 0's blocks
  const int ga = (warp == 0) ? G0 : min(G1, G0 + first + (warp - 1) * per);
  const int gb = (warp == 0) ? G0 + first : min(G1, ga + per);
  const int64_t out_len = (int64_t)(T - 1) * HOP;
  float* ob = out + (int64_t)b * out_len;
  float my_peak = 0.f;
  const bool cold = (warp > 0);
  const int n_head = cold ? min(R - 1, gb - ga) : 0;   // blocks whose sums wait for the previous warp's tail

  float o[16];
#pragma unroll
  for (int r = 0; r < 16; ++r) o[r] = 0.f;

  auto emit = [&](int g) {
    const bool interior = (g >= R - 1) && (g <= T - 1);
    float* op = ob + (int64_t)(g - g_lo) * HOP + lane;
#pragma unroll
    for (int r = 0; r < NR; ++r) {
      float val;
      if (interior) {
        val = o[r] * inv_full[r];
      } else {
        float nrm = 0.f;
#pragma unroll
        for (int q = 0; q < R; ++q) {
          const int tq = g - q;
          if (tq >= 0 && tq <= T - 1) nrm = fmaf(hw[r + NR * q], hw[r + NR * q], nrm);
        }
        val = o[r] / (nrm > 1e-10f ? nrm : 1.0f);
      }
      op[32 * r] = val;
      my_peak = fmaxf(my_peak, fabsf(val));
    }
  };
  // after frame g has been added: close block g (emit it, or park it if it still waits for a tail), shift
  auto close_block = [&](int g) {
    if (g >= ga) {
      if (cold && g - ga < n_head) {
        float* hd = s_head + ((size_t)warp * TAIL + (size_t)(g - ga) * NR) * 32 + lane;
#pragma unroll
        for (int r = 0; r < NR; ++r) hd[32 * r] = o[r];
      } else {
        emit(g);
      }
    }
#pragma unroll
    for (int r = 0; r < TAIL; ++r) o[r] = o[r + NR];
#pragma unroll
    for (int r = TAIL; r < 16; ++r) o[r] = 0.f;
  };

  if (ga < gb) {
    const int t_first = cold ? ga : max(0, ga - (R - 1));
    Window2<HOP> win;
    // KEPT: frames t_first .. t_last of the kept spectrum stream through the ring; frame t lives in stage
    // (t - t_first) % kStages and is the ((t - t_first) / kStages)-th use of that stage's barrier.
    const int t_last = min(gb - 1, T - 1);
    const float4* sp = KEPT ? spec + (spec_utt * T + t_first) * 256 : nullptr;
    // sparse kept spectrum: a frame's block holds only the slots whose noise bit is clear, n_t of them, compacted; n_t
    // comes from the frame's 9 IBM words (lanes 0..8 hold one word each), fetched one frame before its copy is issued
    // (lanes 0..7 hold the frame's 8 ballot words, lane 8 the k-ordered word with the Nyquist bit)
    const uint32_t* bits_b = ibm_bits + (int64_t)b * T * kFW;
    const uint32_t* lane_b = SPARSE ? lane_bits + (int64_t)b * T * 8 : nullptr;
    auto load_word = [&](int f) -> uint32_t {
      if (!SPARSE || f > t_last || lane > 8) return 0u;
      return (lane < 8) ? __ldg(lane_b + (int64_t)f * 8 + lane) : __ldg(bits_b + (int64_t)f * kFW + 8);
    };
    auto count_of = [&](uint32_t wd) -> int {
      if (!SPARSE) return 256;
      int c = (lane < 8) ? __popc(wd) : 0;
      c = __reduce_add_sync(kFull, c);
      const uint32_t b0 = __shfl_sync(kFull, wd, 0) & 1u, b256 = __shfl_sync(kFull, wd, 8) & 1u;
      return 256 - c + (int)(b0 & (b256 ^ 1u));   // slot (0, lane 0) carries DC and Nyquist: kept if either is
    };
    int n_stage[kStages];
    uint32_t wq = 0u;
    if (KEPT) {
#pragma unroll
      for (int st = 0; st < kStages; ++st) {
        n_stage[st] = 0;
        if (t_first + st <= t_last) {
          const int n = count_of(load_word(t_first + st));
          n_stage[st] = n;
          if (lane == 0 && n > 0) {
            mbar_arrive_expect_tx(s_bar + st, 16u * n);
            tma_load_1d(s_ring + st * 256, sp + (size_t)st * 256, 16u * n, s_bar + st);
          }
        }
      }
      wq = load_word(t_first + kStages);
    } else if (t_first <= T - 1) {
      win.load_all(m0, m1, L, t_first, lane);
    }
    // post-filter gains of the frame about to be analysed, fetched one frame ahead
    FrameBits nb;
    LaneBits lq;      // SPARSE: the same bits in (slot, lane) order
    float nmask[9];
    const uint32_t* bw = ibm_bits + ((int64_t)b * T + t_first) * kFW;
    const uint32_t* lw = SPARSE ? lane_bits + ((int64_t)b * T + t_first) * 8 : nullptr;
    const float* mk = mask + (int64_t)b * ml.sb + (int64_t)t_first * ml.st;
    auto fetch_gain = [&]() {
      if (SPARSE) {
        lq.load(lw, bw);
      } else if (gain_mode == GAIN_BITS) {
        nb.load(bw, ln.h);
      } else if (gain_mode != GAIN_NONE) {
#pragma unroll
        for (int j = 0; j < 8; ++j) nmask[j] = __ldg(mk + (int64_t)bin_lo(ln, j) * ml.sf);
        nmask[8] = __ldg(mk + (int64_t)256 * ml.sf);
      }
    };
    if (t_first <= T - 1) fetch_gain();

    float2 Sa[8];
    float ny_a = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) Sa[j] = make_float2(0.f, 0.f);

#pragma unroll 1
    for (int g = t_first; g < gb; ++g) {
      const bool is_first = ((g - t_first) & 1) == 0;
      float2 S[8];
      float s_ny = 0.f;
      if (g <= T - 1) {
        // ---- frame g: forward transform, beamform, post-filter -> S at this lane's low bins (+ Nyquist, lane 0)
        float gj[8], gny;
        if (SPARSE) {
#pragma unroll
          for (int j = 0; j < 8; ++j) gj[j] = lq.bit(j, lane) ? 0.f : 1.f;   // 1 - noise mask
          gny = (lq.ny & 1u) ? 0.f : 1.f;
        } else if (gain_mode == GAIN_BITS) {
#pragma unroll
          for (int j = 0; j < 8; ++j) gj[j] = nb.bit(j, ln.k1) ? 0.f : 1.f;   // 1 - noise mask
          gny = (nb.ny & 1u) ? 0.f : 1.f;
        } else if (gain_mode == GAIN_NONE) {
#pragma unroll
          for (int j = 0; j < 8; ++j) gj[j] = 1.f;
          gny = 1.f;
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j) gj[j] = (gain_mode == GAIN_FLOOR) ? fmaxf(nmask[j], post_floor) : nmask[j];
          gny = (gain_mode == GAIN_FLOOR) ? fmaxf(nmask[8], post_floor) : nmask[8];
        }
        float2 zlo[8], mir[8], zny;
        if (KEPT) {
          const int use = g - t_first, st = use % kStages;
          int n_here = 0;
#pragma unroll
          for (int q = 0; q < kStages; ++q) n_here = (q == st) ? n_stage[q] : n_here;
          if (!SPARSE || n_here > 0) {
            mbar_wait(s_bar + st, (phase_bits >> st) & 1u);
            phase_bits ^= 1u << st;
          }
          const float4* fr = s_ring + st * 256;
          const unsigned lt = (1u << lane) - 1u;
          int base = 0;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const unsigned bal = SPARSE ? lq.keep(j) : kFull;
            const bool keep = (bal >> lane) & 1u;
            const int pos = SPARSE ? base + __popc(bal & lt) : 32 * j + lane;
            base += __popc(bal);
            const float4 q = keep ? fr[pos] : make_float4(0.f, 0.f, 0.f, 0.f);
            zlo[j] = make_float2(q.x, q.y);
            mir[j] = make_float2(q.z, q.w);
          }
          zny = mir[0];                       // lane 0: slot (0).zw is the Nyquist bin ...
          if (lane == 0) mir[0] = zlo[0];     // ... and DC is its own mirror
          __syncwarp();                       // every lane has read the stage: refill it with frame g + kStages
          if (g + kStages <= t_last) {
            const int n = count_of(wq);
#pragma unroll
            for (int q = 0; q < kStages; ++q) n_stage[q] = (q == st) ? n : n_stage[q];
            if (lane == 0 && n > 0) {
              fence_proxy_async();
              mbar_arrive_expect_tx(s_bar + st, 16u * n);
              tma_load_1d(s_ring + st * 256, sp + (size_t)(use + kStages) * 256, 16u * n, s_bar + st);
            }
            wq = load_word(g + kStages + 1);
          }
          if (g + 1 <= T - 1) {
            bw += kFW;
            mk += ml.st;
            if (SPARSE) lw += 8;
            fetch_gain();
          }
        } else {
          if (g + 1 <= T - 1) {
            win.prefetch(m0, m1, L, g + 1, lane);
            bw += kFW;
            mk += ml.st;
            fetch_gain();
          }
          float2 v[16];
          win.frame(v, hw);
#if AVZ_COV_FULLTW
          f512::forward_full(v, sm, ln);
#else
          f512::forward(v, sm, ln);
#endif
          f512::mirror_of_low(v, mir, ln);
#pragma unroll
          for (int j = 0; j < 8; ++j) zlo[j] = v[j];
          zny = v[8];
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 ab = s_ab[bin_lo(ln, j)];
          // a * z + b * conj(m)
          // (z.x, z.y) a.x + (-z.y, z.x) a.y + (b.x, b.y) m.x + (b.y, -b.x) m.y : four packed multiply-adds
          u64 s2 = mul2(pk2(zlo[j]), pk2(ab.x, ab.x));
          s2 = fma2(pk2(-zlo[j].y, zlo[j].x), pk2(ab.y, ab.y), s2);
          s2 = fma2(pk2(ab.z, ab.w), pk2(mir[j].x, mir[j].x), s2);
          s2 = fma2(pk2(ab.w, -ab.z), pk2(mir[j].y, mir[j].y), s2);
          // (a sector pass A did not write - KEEP_SKIP - holds whatever finite value sat there: times gain 0)
          S[j] = up2(mul2(s2, pk2(gj[j], gj[j])));
        }
        const float4 abn = s_ab[256];
        // Re(a z + b conj z), z = Nyquist bin (lane 0)
        s_ny = gny * ((abn.x + abn.z) * zny.x + (abn.w - abn.y) * zny.y);
        if (!KEPT && g + 1 <= T - 1) win.advance();
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) S[j] = make_float2(0.f, 0.f);
      }
      if (is_first && g + 1 < gb) {   // hold it; the next frame shares its inverse transform
#pragma unroll
        for (int j = 0; j < 8; ++j) Sa[j] = S[j];
        ny_a = s_ny;
        continue;
      }
      // ---- one inverse transform for the pair (g-1, g); a lone last frame of the run is the pair (g, nothing)
      if (is_first) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          Sa[j] = S[j];
          S[j] = make_float2(0.f, 0.f);
        }
        ny_a = s_ny;
        s_ny = 0.f;
      }
      const int g_a = is_first ? g : g - 1;
      float2 v[16];
      f512::hermitian_pack(Sa, S, make_float2(ny_a, s_ny), v, ln);
      f512::inverse<(AVZ_APPLY_FULLTW != 0)>(v, sm, ln);
#pragma unroll
      for (int r = 0; r < 16; ++r) o[r] = fmaf(hw[r], v[r].x, o[r]);
      close_block(g_a);
      if (!is_first) {
#pragma unroll
        for (int r = 0; r < 16; ++r) o[r] = fmaf(hw[r], v[r].y, o[r]);
        close_block(g);
      }
    }
  }
  // Blocks t_first .. gb-1 are closed; o[0..TAIL-1] is this run's contribution to the R-1 blocks from gb on:
  // hand it to the next warp (zeros if this warp had no run).
  {
    float* tl = s_tail + (size_t)warp * TAIL * 32 + lane;
#pragma unroll
    for (int r = 0; r < TAIL; ++r) tl[32 * r] = (ga < gb) ? o[r] : 0.f;
  }
  __syncthreads();
  if (cold && ga < gb) {
    // complete this warp's head blocks: own partial sums + the previous warp's tail
    const float* prev = s_tail + (size_t)(warp - 1) * TAIL * 32 + lane;
    const float* hd = s_head + (size_t)warp * TAIL * 32 + lane;
    for (int i = 0; i < n_head; ++i) {
#pragma unroll
      for (int r = 0; r < NR; ++r) o[r] = hd[32 * (i * NR + r)] + prev[32 * (i * NR + r)];
      emit(ga + i);
    }
  }
  if (peak != nullptr) {
    my_peak = warp_max(my_peak);
    if (lane == 0) s_peak[warp] = my_peak;
    __syncthreads();
    float m = 0.f;
#pragma unroll
    for (int i = 0; i < kWarps; ++i) m = fmaxf(m, s_peak[i]);
    if (!cluster_norm) {
      if (threadIdx.x == 0) atomicMax(reinterpret_cast<unsigned int*>(peak + b), __float_as_uint(m));
    } else {
      // Fused peak normalisation (oracle_debug.py:94).  The CTAs of one utterance form a thread-block cluster: they
      // exchange their maxima through distributed shared memory, then every CTA divides the range it has just written
      // (still in L2) - no extra kernel and no second trip of the output to HBM.
      __shared__ float s_cta_peak;
      if (threadIdx.x == 0) s_cta_peak = m;
      cg::cluster_group cl = cg::this_cluster();
      cl.sync();   // every CTA of the utterance has stored its blocks and published its maximum
      float um = 0.f;
      for (unsigned r = 0; r < cl.num_blocks(); ++r) um = fmaxf(um, *cl.map_shared_rank(&s_cta_peak, r));
      cl.sync();   // nobody leaves (and frees its shared memory) while a peer may still be reading it
      if (cl.block_rank() == 0 && threadIdx.x == 0) peak[b] = um;
      const float den = um + peak_eps;
      float4* o4 = reinterpret_cast<float4*>(ob + (int64_t)(G0 - g_lo) * HOP);
      const int n4 = nblk * (HOP / 4);
      constexpr int kStep = kWarps * 32;
      int i = threadIdx.x;
      for (; i + 3 * kStep < n4; i += 4 * kStep) {   // four loads in flight per thread: the pass is L2 latency
        float4 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) v[u] = __ldcg(o4 + i + u * kStep);
#pragma unroll
        for (int u = 0; u < 4; ++u)
          o4[i + u * kStep] = make_float4(__fdiv_rn(v[u].x, den), __fdiv_rn(v[u].y, den), __fdiv_rn(v[u].z, den),
                                          __fdiv_rn(v[u].w, den));
      }
      for (; i < n4; i += kStep) {
        const float4 v = __ldcg(o4 + i);
        o4[i] = make_float4(__fdiv_rn(v.x, den), __fdiv_rn(v.y, den), __fdiv_rn(v.z, den), __fdiv_rn(v.w, den));
      }
    }
  }
}

template <int HOP, bool KEPT, bool SPARSE>
__global__ void __launch_bounds__(kWarps * 32, KEPT ? AVZ_MINB_APPLY_KEPT : AVZ_MINB_APPLY)
k512_apply(const float* __restrict__ mix, const float4* __restrict__ spec, const float2* __restrict__ wgt,
           const uint32_t* __restrict__ ibm_bits, const float* __restrict__ mask, MaskLayout ml, int gain_mode,
           float post_floor, int L, int T, int blocks_per_cta, float* __restrict__ out, float* __restrict__ peak,
           int cluster_norm, float peak_eps, const uint32_t* __restrict__ shdr, const uint32_t* __restrict__ lane_bits,
           Tables tb) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  uint32_t phase_bits = 0u;
  apply_body<HOP, KEPT, SPARSE>(smem_raw, blockIdx.y, blockIdx.x, true, phase_bits, blockIdx.y, mix, spec, wgt, ibm_bits, mask,
                                ml, gain_mode, post_floor, L, T, blocks_per_cta, out, peak, cluster_norm, peak_eps, shdr,
                                lane_bits, tb);
}

// ------------------------------------------------------------------------------------------
// Pass A + weights + pass B + normalisation as ONE persistent kernel (oracle-IBM path)
// ------------------------------------------------------------------------------------------
// Measured (profiles/README.md, "what bounds pass A / pass B"): with the spectrum handed from pass A to pass B through
// HBM, k512_cov is bound by its 2 GB store stream and k512_apply by the 2 GB read-back (70-76 % of the DRAM peak), not
// by instructions.  Here the two passes are tasks of one launch, dequeued in an order that lets pass B of an utterance
// run a few microseconds after its pass A, and the kept spectrum lives in a RING of utterance slots small enough to
// stay in the 126 MB L2: a slot is overwritten while its lines are still dirty in L2, so the spectrum never reaches
// DRAM in either direction.  Finalize, the 2x2 solve and the peak normalisation ride along (the CTA that completes an
// utterance's last pass-A / pass-B task does them), which also removes three launches.
//
// Task queue (one atomic counter; tasks are dequeued in index order by whichever CTA is free):
//   slot s = 0, 1, ...:  [A(s, chunk 0..CA-1)]  then  [B(s - LAG, chunk 0..CB-1)]
// A(u, c) first waits until B(u - NSLOT) is complete (its ring slot is free); B(u, c) waits until the weights of u are
// published.  Every wait targets tasks with a SMALLER queue index; those were dequeued earlier by CTAs that are
// running, and the induction closes on task 0, which waits for nothing - so the kernel cannot deadlock whatever the
// number of resident CTAs.  Results are bit-identical to the separate kernels (same device functions, same order).
struct FusedArgs {
  const float* mix;
  const uint32_t* ibm_bits;
  const float2* dvec;
  float* part;       // [B][CA][5][kFP]
  float4* spec;      // [NSLOT][T][256]
  float4* R;         // [B][F]
  float* msum;       // [B][F]
  float2* w;         // [B][F][2]
  float* out;        // [B][(T-1) HOP]
  float* peak;       // [B]
  int* ctrl;         // [0] queue head, then a_done[B], w_ready[B], b_done[B]
  AvzMvdrCfg cfg;
  float norm_eps, peak_eps;   // peak_eps < 0: no normalisation
  int B, L, T, fpt, CA, CB, lag, nslot;
  Tables tb;
};

template <int HOP>
__host__ __device__ constexpr size_t fused_smem_bytes() {
  constexpr size_t cov = (size_t)kWarps * f512::kSmemComplex * sizeof(float2) + (size_t)kWarps * 5 * kFP * sizeof(float);
  constexpr size_t app = apply_smem_bytes<HOP, true>();
  return cov > app ? cov : app;
}

__device__ __forceinline__ void spin_until(const int* flag, int want, int what, int u) {
  if (threadIdx.x == 0) {
    unsigned spins = 0;
    while (ld_acquire_gpu(flag) < want) {
      __nanosleep(64);
      if (++spins > (1u << 19)) {   // ~a second: a protocol error must not hang the device
        printf("k512_fused: wait %d on utterance %d timed out (flag %d < %d, block %d)\n", what, u, ld_acquire_gpu(flag), want,
               (int)blockIdx.x);
        __trap();
      }
    }
  }
  __syncthreads();
}

template <int HOP>
__global__ void __launch_bounds__(kWarps * 32, 2) k512_fused(FusedArgs a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ int s_task, s_last;
  int* queue = a.ctrl;
  int* a_done = a.ctrl + 1;
  int* w_ready = a_done + a.B;
  int* b_done = w_ready + a.B;
  const int lag = a.lag;
  const int head = lag * a.CA;                       // slots 0..lag-1: pass A only
  const int mid = (a.B - lag) * (a.CA + a.CB);       // slots lag..B-1: A(s) then B(s - lag)
  const int n_tasks = a.B * (a.CA + a.CB);
  const int64_t out_len = (int64_t)(a.T - 1) * HOP;
  bool ring_used = false;
  uint32_t phase_bits = 0u;      // per warp: parity of the next wait on each stage of its TMA ring
  for (;;) {
    if (threadIdx.x == 0) s_task = atomicAdd(queue, 1);
    __syncthreads();
    const int t = s_task;
    if (t >= n_tasks) break;
    bool is_a;
    int u, c;
    if (t < head) {
      is_a = true, u = t / a.CA, c = t - u * a.CA;
    } else if (t < head + mid) {
      const int q = t - head, slot = q / (a.CA + a.CB), r = q - slot * (a.CA + a.CB);
      if (r < a.CA) is_a = true, u = lag + slot, c = r;
      else is_a = false, u = slot, c = r - a.CA;
    } else {
      const int q = t - head - mid;
      is_a = false, u = (a.B - lag) + q / a.CB, c = q % a.CB;
    }
    const int ring = u % a.nslot;
    if (is_a) {
      if (u >= a.nslot) spin_until(b_done + (u - a.nslot), a.CB, 0, u);   // the ring slot's previous utterance has been consumed
      cov_body<HOP, W_BITS, true, KEEP_ALL>(smem_raw, u, c, a.CA, a.mix, a.ibm_bits, nullptr, MaskLayout{0, 0, 0, nullptr, 0u, 0u},
                                         a.L, a.T, a.fpt, 0.f, a.part, a.spec, ring, nullptr, a.tb);
      __threadfence();
      __syncthreads();
      if (threadIdx.x == 0) s_last = (atomicAdd(a_done + u, 1) == a.CA - 1);
      __syncthreads();
      if (s_last) {
        // last pass-A task of utterance u: float64 sum of the chunk partials in k_cov_finalize's order (chunks c = q mod 4
        // summed per slice q, slices combined in order), then the closed-form weights of k_mvdr_weights
        __threadfence();
        finalize_weights_utt(a.part, u, a.CA, a.norm_eps, a.dvec, a.cfg, a.R, a.msum, a.w);
        __threadfence();
        __syncthreads();
        if (threadIdx.x == 0) st_release_gpu(w_ready + u, 1);
      }
    } else {
      spin_until(w_ready + u, 1, 1, u);
      fence_proxy_async_all();   // the spectrum was written by other CTAs' ordinary stores; it is read by TMA bulk copies here
      apply_body<HOP, true, false>(smem_raw, u, c, !ring_used, phase_bits, ring, nullptr, a.spec, a.w, a.ibm_bits, nullptr,
                            MaskLayout{0, 0, 0, nullptr, 0u, 0u}, a.cfg.post_mode == AVZ_POST_ONE_MINUS_NOISE ? GAIN_BITS : GAIN_NONE,
                            0.f, a.L, a.T, a.fpt, a.out, a.peak, 0, 0.f, nullptr, nullptr, a.tb);
      ring_used = true;
      __threadfence();
      __syncthreads();
      if (threadIdx.x == 0) s_last = (atomicAdd(b_done + u, 1) == a.CB - 1);
      __syncthreads();
      if (s_last && a.peak_eps >= 0.f) {
        // last pass-B task of utterance u: x / (max|x| + eps) while the output is still in L2 (k_peak_normalise's arithmetic)
        __threadfence();
        const float den = __ldcg(a.peak + u) + a.peak_eps;
        float* xb = a.out + (int64_t)u * out_len;
        float4* x4 = reinterpret_cast<float4*>(xb);
        const int n4 = (int)(out_len >> 2);           // HOP is a multiple of 4 and the row starts 16-byte aligned
        constexpr int kStep = kWarps * 32;
        int i = threadIdx.x;
        for (; i + 3 * kStep < n4; i += 4 * kStep) {
          float4 v[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) v[q] = __ldcg(x4 + i + q * kStep);
#pragma unroll
          for (int q = 0; q < 4; ++q)
            x4[i + q * kStep] = make_float4(__fdiv_rn(v[q].x, den), __fdiv_rn(v[q].y, den), __fdiv_rn(v[q].z, den),
                                            __fdiv_rn(v[q].w, den));
        }
        for (; i < n4; i += kStep) {
          const float4 v = __ldcg(x4 + i);
          x4[i] = make_float4(__fdiv_rn(v.x, den), __fdiv_rn(v.y, den), __fdiv_rn(v.z, den), __fdiv_rn(v.w, den));
        }
      }
    }
    __syncthreads();   // shared memory (task id, transposition buffers, ring) is reused by the next task
  }
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
static int ctas_per_sm_target() {
  // CTAs per SM a launch should have at least, so that the tail of the last wave is short (AVZ_CTAS_PER_SM overrides)
  // (it decides how an utterance is cut into chunks, i.e. the summation order of the covariance partials: fixed in
  // the release library, so that results depend on the arguments only)
#ifdef AVZ_EXPERIMENT
  static const int v = [] {
    int t = 20;
    if (const char* e = getenv("AVZ_CTAS_PER_SM")) t = atoi(e);
    return t < 1 ? 1 : t;
  }();
  return v;
#else
  return 20;
#endif
}

static int frames_per_cta(int B, int T, int sms) {
  int per_utt = (ctas_per_sm_target() * sms + B - 1) / B;
  if (per_utt < 1) per_utt = 1;
  int fpc = (T + per_utt - 1) / per_utt;
  if (fpc < 8 * kWarps) fpc = 8 * kWarps;
  if (fpc > T) fpc = T;
  return fpc;
}

int cov_chunks512(int B, int T) {
  const int fpc = frames_per_cta(B, T, num_sms());
  return (T + fpc - 1) / fpc;
}

// workspace layout: [partial sums: B * chunks * 5 * kFP floats][pad to 16][count: 4 x u32][entries: cap x u64]
static size_t part_bytes(int B, int chunks) {
  return (((size_t)B * chunks * 5 * kFP * sizeof(float)) + 15) / 16 * 16;
}
static unsigned amb_cap(int B, int T) {
  const unsigned long long c = (unsigned long long)B * T * 8ull;   // 8 near-ties per frame on average
  return (unsigned)(c > (1ull << 26) ? (1ull << 26) : c);
}
int64_t ws_bytes512(int B, int T) {   // partials, near-tie counter + list, per-utterance completion counters (k512_cov_w)
  return (int64_t)(part_bytes(B, cov_chunks512(B, T)) + 16 + (size_t)amb_cap(B, T) * 8 + (size_t)B * sizeof(int));
}

// kept-spectrum buffer: [B*T frames x 4096 B][per-utterance completion counters, padded to 256 B][transposed mask]
static size_t spec_hdr_offset(int B, int T) {
  return (size_t)B * T * 4096 + (((size_t)B * sizeof(unsigned int)) + 255) / 256 * 256;
}
static size_t spec_mask_offset(int B, int T) { return spec_hdr_offset(B, T) + kMaskHdrBytes; }
static uint32_t* spec_hdr(const void* spec, int B, int T) {
  return reinterpret_cast<uint32_t*>(const_cast<unsigned char*>(static_cast<const unsigned char*>(spec)) + spec_hdr_offset(B, T));
}
int64_t spec_ws_bytes512(int B, int T) {
  return (int64_t)(spec_mask_offset(B, T) + (size_t)B * T * kMaskPitch * sizeof(float));
}
// Mask reads of the fused kernels: transposed copy behind the kept spectrum when there is one, else the caller's.
static const float* stage_mask(const float* mask, const void* spec, int B, int T, MaskLayout* ml, cudaStream_t st) {
  if (mask == nullptr || spec == nullptr) {
    *ml = MaskLayout{(int64_t)kF * T, T, 1, nullptr, 0u, 0u};
    return mask;
  }
  float* mask_t = reinterpret_cast<float*>(const_cast<unsigned char*>(static_cast<const unsigned char*>(spec)) +
                                           spec_mask_offset(B, T));
  k_mask_transpose<<<dim3((T + 31) / 32, (kMaskPitch + 31) / 32, B), 256, 0, st>>>(mask, mask_t, T, spec_hdr(spec, B, T));
  *ml = MaskLayout{(int64_t)T * kMaskPitch, 1, kMaskPitch, nullptr, 0u, 0u};
  return mask_t;
}

static float ibm_tol2() {
  // (float32 FFT error bound of one bin)^2, relative to the rms bin magnitude of the frame.  The error of this
  // transform is heavy-tailed, not Gaussian: deciding EVERY bin of a full config-5 job (65 536 utterances, 8.4e9 bins)
  // in float64 (avz_ibm_exact_f32, tools/ibm_exact_c5.py) found 19 wrong float32 decisions outside a 5e-7 band (the
  // round-1 value, chosen from 33 M bins) and none outside 1e-6, 2e-6 or 4e-6 (profiles/r2_ibm_exact_c5.json).  The
  // shipped bound is 4e-6 - the generic path's value, 4-8x above the first observed failure - at +0.08 ms of fix-up per
  // 1024 x 4 s.  Fixed at compile time: only builds with -DAVZ_EXPERIMENT read the AVZ_IBM_TOL environment variable
  // (tools/ibm_tol_scan.py), so a stray variable cannot change results.
  static const float t2 = [] {
    double tol = 4e-6;
#ifdef AVZ_EXPERIMENT
    if (const char* e = getenv("AVZ_IBM_TOL")) tol = atof(e);
#endif
    return (float)(tol * tol);
  }();
  return t2;
}

// tail (optional, IBM path only): fold k_cov_finalize + k_mvdr_weights into pass A (k512_cov_w); its counters live behind
// the near-tie list in the workspace
template <int HOP>
int launch_ibm_cov(const float* mix, const float* tgt, const float* itf, const float* mask, int B, int64_t L,
                   float sqrt_eps, uint32_t* ibm_bits, float* part, int* chunks_out, void* spec, cudaStream_t st,
                   const CovTailArgs* tail, int sparse) {
  Tables tb;
  int rc = tables_for(kN, &tb);
  if (rc) return rc;
  if (L >= (1ll << 30)) return set_error(AVZ_EINVAL, "L=%lld too long for the 512-point fast path", (long long)L);
  const int T = (int)avz_num_frames(L, kN, HOP);
  const int fpc = frames_per_cta(B, T, num_sms());
  const int chunks = (T + fpc - 1) / fpc;
  *chunks_out = chunks;
  dim3 grid(chunks, B);
  const size_t smem_fft = (size_t)kWarps * f512::kSmemComplex * sizeof(float2);
  if (mask == nullptr) {
    unsigned char* wsb = reinterpret_cast<unsigned char*>(part) + part_bytes(B, chunks);
    AmbList al;
    al.count = reinterpret_cast<unsigned int*>(wsb);
    al.entries = reinterpret_cast<unsigned long long*>(wsb + 16);
    al.cap = amb_cap(B, T);
    AVZ_CUDA_OK(cudaMemsetAsync(al.count, 0, 16, st));
    if (tail != nullptr) AVZ_CUDA_OK(cudaMemsetAsync(wsb + 16 + (size_t)al.cap * 8, 0, (size_t)B * sizeof(int), st));
    if (spec != nullptr)   // this pass A stages no mask: a header left by an earlier learned-mask call must not survive
      AVZ_CUDA_OK(cudaMemsetAsync(spec_hdr(spec, B, T), 0, 16, st));
    AVZ_CUDA_OK(cudaFuncSetAttribute(k512_ibm<HOP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_fft));
    prof_begin(PROF_IBM, st);
    // sparse kept spectrum: the bits are also kept in (slot, lane) order, in the (unused) mask region of `spec`
    uint32_t* lane_bits = (spec && sparse == KEEP_SPARSE) ? reinterpret_cast<uint32_t*>(static_cast<unsigned char*>(spec) + spec_mask_offset(B, T))
                                           : nullptr;
    k512_ibm<HOP><<<grid, kWarps * 32, smem_fft, st>>>(tgt, itf, (int)L, T, fpc, ibm_bits, lane_bits, al, ibm_tol2(), tb);
    prof_end(PROF_IBM, st);
    AVZ_LAUNCH_OK("k512_ibm");
    prof_begin(PROF_FIXUP, st);
    k512_ibm_fixup<<<num_sms() * 32, 256, 0, st>>>(tgt, itf, L, T, HOP, B, ibm_bits, lane_bits, al, tb);
    prof_end(PROF_FIXUP, st);
    AVZ_LAUNCH_OK("k512_ibm_fixup");
  }
  const uint32_t* lane_bits_c = (spec && sparse == KEEP_SPARSE && mask == nullptr)
                                    ? reinterpret_cast<const uint32_t*>(static_cast<unsigned char*>(spec) + spec_mask_offset(B, T))
                                    : nullptr;
  const size_t smem_cov = smem_fft + (size_t)kWarps * 5 * kFP * sizeof(float);
  prof_begin(PROF_COV, st);
  if (mask == nullptr && tail != nullptr) {
    unsigned char* wsb = reinterpret_cast<unsigned char*>(part) + part_bytes(B, chunks);
    CovTail ct;
    ct.done = reinterpret_cast<int*>(wsb + 16 + (size_t)amb_cap(B, T) * 8);
    ct.dvec = reinterpret_cast<const float2*>(tail->dvec);
    ct.R = reinterpret_cast<float4*>(tail->R);
    ct.msum = tail->msum;
    ct.w = reinterpret_cast<float2*>(tail->w);
    ct.cfg = *tail->cfg;
    ct.norm_eps = tail->norm_eps;
    ct.shdr = spec ? spec_hdr(spec, B, T) : nullptr;
    ct.sparse = spec ? sparse : 0;
    ct.lane_bits = lane_bits_c;
    auto kern = (ct.sparse == KEEP_SPARSE) ? k512_cov_w<HOP, KEEP_SPARSE>
              : (ct.sparse == KEEP_SKIP)   ? k512_cov_w<HOP, KEEP_SKIP> : k512_cov_w<HOP, KEEP_ALL>;
    AVZ_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_cov));
    kern<<<grid, kWarps * 32, smem_cov, st>>>(mix, ibm_bits, (int)L, T, fpc, part, reinterpret_cast<float4*>(spec), ct, tb);
  } else if (mask == nullptr) {
    uint32_t* shdr = spec ? spec_hdr(spec, B, T) : nullptr;
    const int km = spec ? sparse : 0;
    auto kern = (km == KEEP_SPARSE) ? k512_cov<HOP, W_BITS, KEEP_SPARSE>
              : (km == KEEP_SKIP)   ? k512_cov<HOP, W_BITS, KEEP_SKIP> : k512_cov<HOP, W_BITS, KEEP_ALL>;
    AVZ_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_cov));
    kern<<<grid, kWarps * 32, smem_cov, st>>>(mix, ibm_bits, nullptr, MaskLayout{0, 0, 0, nullptr, 0u, 0u}, (int)L, T, fpc, 0.f,
                                              part, reinterpret_cast<float4*>(spec), shdr, lane_bits_c, tb);
  } else {
    AVZ_CUDA_OK(cudaFuncSetAttribute(k512_cov<HOP, W_MASK, KEEP_ALL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_cov));
    MaskLayout ml;
    const float* mptr = stage_mask(mask, spec, B, T, &ml, st);
    AVZ_LAUNCH_OK("k_mask_transpose");
    k512_cov<HOP, W_MASK, KEEP_ALL><<<grid, kWarps * 32, smem_cov, st>>>(mix, nullptr, mptr, ml, (int)L, T, fpc, sqrt_eps, part,
                                                                      reinterpret_cast<float4*>(spec), nullptr, nullptr, tb);
  }
  prof_end(PROF_COV, st);
  AVZ_LAUNCH_OK("k512_cov");
  return AVZ_OK;
}

template <int HOP>
int launch_apply(const float* mix, const void* spec, const float* w, const uint32_t* ibm_bits, const float* mask,
                 int gain_mode, float post_floor, int B, int64_t L, float* out, float* peak, int fuse_norm,
                 float peak_eps, int mask_staged, cudaStream_t st, int sparse) {
  Tables tb;
  int rc = tables_for(kN, &tb);
  if (rc) return rc;
  if (L >= (1ll << 30)) return set_error(AVZ_EINVAL, "L=%lld too long for the 512-point fast path", (long long)L);
  const int T = (int)avz_num_frames(L, kN, HOP);
  const int n_blocks = T - 1;
  if (n_blocks <= 0) return AVZ_OK;
  int bpc = frames_per_cta(B, n_blocks, num_sms());
  const int chunks = (n_blocks + bpc - 1) / bpc;
  const size_t smem = (spec != nullptr) ? apply_smem_bytes<HOP, true>() : apply_smem_bytes<HOP, false>();
  dim3 grid(chunks, B);
  MaskLayout ml;
  const float* mptr;
  if (mask_staged && spec != nullptr) {
    // the transposed copy the learned-mask pass A left behind the kept spectrum: no second transposition
    ml = MaskLayout{(int64_t)T * kMaskPitch, 1, kMaskPitch, spec_hdr(spec, B, T), (uint32_t)B, (uint32_t)T};
    mptr = reinterpret_cast<const float*>(static_cast<const unsigned char*>(spec) + spec_mask_offset(B, T));
  } else {
    mptr = stage_mask((gain_mode == GAIN_FLOOR || gain_mode == GAIN_MASK) ? mask : nullptr, spec, B, T, &ml, st);
    AVZ_LAUNCH_OK("k_mask_transpose");
  }
  prof_begin(PROF_APPLY, st);
  if (spec != nullptr) {
    const bool sp_on = sparse == KEEP_SPARSE && gain_mode == GAIN_BITS;
    auto kern = sp_on ? k512_apply<HOP, true, true> : k512_apply<HOP, true, false>;
    AVZ_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    // fused normalisation: the CTAs of an utterance run as one thread-block cluster (portable size limit 8)
    const int cluster_norm = (fuse_norm && peak != nullptr && chunks <= 8) ? 1 : 0;
    cudaLaunchConfig_t lc = {};
    lc.gridDim = grid;
    lc.blockDim = dim3(kWarps * 32);
    lc.dynamicSmemBytes = smem;
    lc.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = cluster_norm ? chunks : 1;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    lc.attrs = at;
    lc.numAttrs = 1;
    // the header word that records the layout is written by the IBM pass A only: a float-mask pass B does not look at it
    const uint32_t* shdr = spec_hdr(spec, B, T);
    AVZ_CUDA_OK(cudaLaunchKernelEx(&lc, kern, (const float*)nullptr, reinterpret_cast<const float4*>(spec),
                                   reinterpret_cast<const float2*>(w), ibm_bits, mptr, ml, gain_mode, post_floor, (int)L, T,
                                   bpc, out, peak, cluster_norm, peak_eps, shdr,
                                   sp_on ? reinterpret_cast<const uint32_t*>(static_cast<const unsigned char*>(spec) +
                                                                             spec_mask_offset(B, T))
                                         : (const uint32_t*)nullptr,
                                   tb));
    if (fuse_norm && !cluster_norm) {   // too many chunks per utterance for a cluster: separate pass
      prof_end(PROF_APPLY, st);
      AVZ_LAUNCH_OK("k512_apply");
      return avz_peak_normalise_f32(out, B, (int64_t)(T - 1) * HOP, peak, peak_eps, st);
    }
  } else {
    AVZ_CUDA_OK(cudaFuncSetAttribute(k512_apply<HOP, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k512_apply<HOP, false, false><<<grid, kWarps * 32, smem, st>>>(mix, nullptr, reinterpret_cast<const float2*>(w), ibm_bits,
                                                                   mptr, ml, gain_mode, post_floor, (int)L, T, bpc, out, peak,
                                                                   0, 0.f, (const uint32_t*)nullptr, (const uint32_t*)nullptr, tb);
  }
  prof_end(PROF_APPLY, st);
  AVZ_LAUNCH_OK("k512_apply");
  return AVZ_OK;
}

// ---- fused oracle path: k512_ibm + k512_ibm_fixup, then k512_fused (pass A, weights, pass B, normalisation)
static int fused_env(const char* name, int dflt) {
#ifdef AVZ_EXPERIMENT
  if (const char* e = getenv(name)) return atoi(e);
#endif
  (void)name;
  return dflt;
}
struct FusedGeo {
  int T, fpt, CA, CB, lag, nslot;
  size_t off_amb, off_ctrl, off_spec, total;
};
static FusedGeo fused_geo(int B, int64_t L, int hop) {
  FusedGeo g;
  g.T = (int)avz_num_frames(L, kN, hop);
  g.fpt = fused_env("AVZ_FUSED_FPT", 32);            // frames (pass A) / hop-blocks (pass B) per task
  if (g.fpt < 8 * kWarps) g.fpt = 8 * kWarps;
  g.CA = (g.T + g.fpt - 1) / g.fpt;
  g.CB = (g.T - 1 + g.fpt - 1) / g.fpt;
  // pass B of an utterance is dequeued `lag` utterances after its pass A: far enough behind that the ~2 x SMs tasks in
  // flight have retired its pass A (no waiting), close enough that lag + in-flight utterances of spectrum fit in L2
  const int inflight = (2 * num_sms() + g.CA + g.CB - 1) / (g.CA + g.CB);
  g.lag = fused_env("AVZ_FUSED_LAG", inflight + 2);
  if (g.lag > B) g.lag = B;
  if (g.lag < 1) g.lag = 1;
  g.nslot = fused_env("AVZ_FUSED_NSLOT", g.lag + inflight + 2);
  if (g.nslot > B) g.nslot = B;
  if (g.nslot < g.lag + 1 && g.nslot < B) g.nslot = g.lag + 1;
  g.off_amb = (((size_t)B * g.CA * 5 * kFP * sizeof(float)) + 15) / 16 * 16;
  g.off_ctrl = g.off_amb + 16 + (size_t)amb_cap(B, g.T) * 8;
  g.off_spec = (g.off_ctrl + (size_t)(1 + 3 * (size_t)B) * sizeof(int) + 255) / 256 * 256;
  g.total = g.off_spec + (size_t)g.nslot * g.T * 4096;
  return g;
}
int64_t fused_ws_bytes(int B, int64_t L, int hop) { return (int64_t)fused_geo(B, L, hop).total; }

template <int HOP>
int launch_oracle_fused(const float* mix, const float* tgt, const float* itf, int B, int64_t L, const AvzMvdrCfg* cfg,
                        float norm_eps, float peak_eps, const float* dvec, uint32_t* ibm_bits, float* R, float* msum,
                        float* w, float* out, float* peak, void* ws, cudaStream_t st) {
  Tables tb;
  int rc = tables_for(kN, &tb);
  if (rc) return rc;
  if (L >= (1ll << 30)) return set_error(AVZ_EINVAL, "L=%lld too long for the 512-point fast path", (long long)L);
  const FusedGeo g = fused_geo(B, L, HOP);
  if (g.T < 2) return set_error(AVZ_EINVAL, "signal too short");
  unsigned char* wsb = static_cast<unsigned char*>(ws);
  AmbList al;
  al.count = reinterpret_cast<unsigned int*>(wsb + g.off_amb);
  al.entries = reinterpret_cast<unsigned long long*>(wsb + g.off_amb + 16);
  al.cap = amb_cap(B, g.T);
  AVZ_CUDA_OK(cudaMemsetAsync(al.count, 0, 16, st));
  AVZ_CUDA_OK(cudaMemsetAsync(wsb + g.off_ctrl, 0, (size_t)(1 + 3 * (size_t)B) * sizeof(int), st));
  AVZ_CUDA_OK(cudaMemsetAsync(peak, 0, (size_t)B * sizeof(float), st));
  const size_t smem_fft = (size_t)kWarps * f512::kSmemComplex * sizeof(float2);
  const int fpc = frames_per_cta(B, g.T, num_sms());
  AVZ_CUDA_OK(cudaFuncSetAttribute(k512_ibm<HOP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_fft));
  prof_begin(PROF_IBM, st);
  k512_ibm<HOP><<<dim3((g.T + fpc - 1) / fpc, B), kWarps * 32, smem_fft, st>>>(tgt, itf, (int)L, g.T, fpc, ibm_bits, nullptr, al,
                                                                                ibm_tol2(), tb);
  prof_end(PROF_IBM, st);
  AVZ_LAUNCH_OK("k512_ibm");
  prof_begin(PROF_FIXUP, st);
  k512_ibm_fixup<<<num_sms() * 32, 256, 0, st>>>(tgt, itf, L, g.T, HOP, B, ibm_bits, nullptr, al, tb);
  prof_end(PROF_FIXUP, st);
  AVZ_LAUNCH_OK("k512_ibm_fixup");
  FusedArgs fa;
  fa.mix = mix;
  fa.ibm_bits = ibm_bits;
  fa.dvec = reinterpret_cast<const float2*>(dvec);
  fa.part = reinterpret_cast<float*>(wsb);
  fa.spec = reinterpret_cast<float4*>(wsb + g.off_spec);
  fa.R = reinterpret_cast<float4*>(R);
  fa.msum = msum;
  fa.w = reinterpret_cast<float2*>(w);
  fa.out = out;
  fa.peak = peak;
  fa.ctrl = reinterpret_cast<int*>(wsb + g.off_ctrl);
  fa.cfg = *cfg;
  fa.norm_eps = norm_eps;
  fa.peak_eps = peak_eps;
  fa.B = B;
  fa.L = (int)L;
  fa.T = g.T;
  fa.fpt = g.fpt;
  fa.CA = g.CA;
  fa.CB = g.CB;
  fa.lag = g.lag;
  fa.nslot = g.nslot;
  fa.tb = tb;
  const size_t smem = fused_smem_bytes<HOP>();
  AVZ_CUDA_OK(cudaFuncSetAttribute(k512_fused<HOP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int per_sm = 0;
  AVZ_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k512_fused<HOP>, kWarps * 32, smem));
  if (per_sm < 1) return set_error(AVZ_ECUDA, "k512_fused does not fit on this device");
  int64_t grid = (int64_t)per_sm * num_sms();
  const int64_t n_tasks = (int64_t)B * (g.CA + g.CB);
  if (grid > n_tasks) grid = n_tasks;
  prof_begin(PROF_COV, st);
  k512_fused<HOP><<<(unsigned)grid, kWarps * 32, smem, st>>>(fa);
  prof_end(PROF_COV, st);
  AVZ_LAUNCH_OK("k512_fused");
  return AVZ_OK;
}
template int launch_oracle_fused<128>(const float*, const float*, const float*, int, int64_t, const AvzMvdrCfg*, float, float,
                                      const float*, uint32_t*, float*, float*, float*, float*, float*, void*, cudaStream_t);
template int launch_oracle_fused<256>(const float*, const float*, const float*, int, int64_t, const AvzMvdrCfg*, float, float,
                                      const float*, uint32_t*, float*, float*, float*, float*, float*, void*, cudaStream_t);

// Every bin decided in float64 (slow; the checker for the float32 + fix-up path at sizes the CPU oracle cannot reach).
int launch_ibm_exact(const float* tgt, const float* itf, int B, int64_t L, int hop, uint32_t* ibm_bits, void* ws16,
                     cudaStream_t st) {
  Tables tb;
  int rc = tables_for(kN, &tb);
  if (rc) return rc;
  const int T = (int)avz_num_frames(L, kN, hop);
  AVZ_CUDA_OK(cudaMemsetAsync(ibm_bits, 0, (size_t)B * T * kFW * sizeof(uint32_t), st));
  AVZ_CUDA_OK(cudaMemsetAsync(ws16, 0xff, 16, st));   // count = 0xffffffff > cap = 0: "overflow" -> scan all bins
  AmbList al;
  al.count = reinterpret_cast<unsigned int*>(ws16);
  al.entries = nullptr;
  al.cap = 0;
  k512_ibm_fixup<<<num_sms() * 8, 256, 0, st>>>(tgt, itf, L, T, hop, B, ibm_bits, nullptr, al, tb);
  AVZ_LAUNCH_OK("k512_ibm_fixup(exact)");
  return AVZ_OK;
}

template int launch_ibm_cov<128>(const float*, const float*, const float*, const float*, int, int64_t, float, uint32_t*,
                                 float*, int*, void*, cudaStream_t, const CovTailArgs*, int);
template int launch_ibm_cov<256>(const float*, const float*, const float*, const float*, int, int64_t, float, uint32_t*,
                                 float*, int*, void*, cudaStream_t, const CovTailArgs*, int);
template int launch_apply<128>(const float*, const void*, const float*, const uint32_t*, const float*, int, float, int,
                               int64_t, float*, float*, int, float, int, cudaStream_t, int);
template int launch_apply<256>(const float*, const void*, const float*, const uint32_t*, const float*, int, float, int,
                               int64_t, float*, float*, int, float, int, cudaStream_t, int);

}  // namespace o512
}  // namespace avz

// ------------------------------------------------------------------------------------------
// Streaming step (BASELINE config 4).  NOT IN THE REFERENCE - this project's definition:
//   R_t = lam R_{t-1} + (1 - lam) m_t y_t y_t^H ,  n_t = lam n_{t-1} + (1 - lam) m_t
//   w_t = mvdr(R_t / (n_t + norm_eps) + sigma I) ,  S_t = w_t^H y_t  (bins below the high-pass -> 0 / mic 0)
// with the same 512 / 128 STFT framing and overlap-add as the batch path, one hop of 128 new samples per call.
// One warp per stream.  Per stream state: the last 384 input samples of both mics, the 384 open output samples,
// and (R00, R11, Re R01, Im R01, n) per bin in the lane layout of the transform.
// ------------------------------------------------------------------------------------------
namespace avz {
namespace o512 {

constexpr int kStreamStateFloats = 2 * 384 + 384 + 5 * 288;   // hist, tail, cov (8 x 32 + Nyquist row, padded to 288)

// The stream's whole state (10 KB) and its new hop are brought into shared memory by two TMA bulk copies issued at
// kernel entry, so the ~600 ns trip to L2/HBM overlaps the per-lane set-up instead of stalling every dependent load
// (the kernel is one long dependent chain per warp; ncu before this change: 3 long-scoreboard stalls per issue).
constexpr int kStreamStageFloats = kStreamStateFloats + 256;   // state + new hop [2][128]
// Per warp: [new hop 256 floats][history 768][open tail 384][covariance 1440].  The first 1408 floats are copied
// into registers before the transform starts, so the transform's transposition buffer (5376 B) reuses them: 45.6 KB
// per CTA, four CTAs per SM.
constexpr size_t kStreamSmem = (size_t)kWarps * kStreamStageFloats * sizeof(float) + (size_t)kWarps * sizeof(uint64_t);
static_assert((256 + 768 + 384) * sizeof(float) >= f512::kSmemComplex * sizeof(float2), "transposition buffer must fit");

// MVDR weights of one bin from the recursive statistics, float64, one division.  With nne = n + norm_eps,
// A' = R00 + sigma nne, C' = R11 + sigma nne, b' = R01 (all "times nne"), det' = A' C' - |b'|^2, u' = adj(.) d:
//   w = u / (d^H u + w_eps),  u = (R/nne + sigma I)^-1 d = u' nne / det'   =>   w = u' nne / (d^H u' nne + w_eps det')
// An exactly singular (or non-finite) system gives w = [1, 0] like the batch kernel (masked_mvdr.py:120-122).
__device__ __forceinline__ void stream_weights(float r00, float r11, float rre, float rim, float nn, float2 d0f,
                                               float2 d1f, const AvzMvdrCfg& cfg, float2& w0, float2& w1) {
  const double nne = (double)nn + (double)cfg.norm_eps;
  const double sg = (double)cfg.sigma * nne;
  const double aa = (double)r00 + sg, cc = (double)r11 + sg, bx = (double)rre, by = (double)rim;
  const double d0x = (double)d0f.x, d0y = (double)d0f.y, d1x = (double)d1f.x, d1y = (double)d1f.y;
  const double det = aa * cc - (bx * bx + by * by);
  w0 = make_float2(1.f, 0.f);
  w1 = make_float2(0.f, 0.f);
  if (det == 0.0 || !isfinite(det)) return;
  // u0' = C' d0 - b' d1,  u1' = A' d1 - conj(b') d0
  const double u0x = cc * d0x - (bx * d1x - by * d1y), u0y = cc * d0y - (bx * d1y + by * d1x);
  const double u1x = aa * d1x - (bx * d0x + by * d0y), u1y = aa * d1y - (bx * d0y - by * d0x);
  // D = nne (conj(d0) u0' + conj(d1) u1') + w_eps det'
  const double dx = nne * (d0x * u0x + d0y * u0y + d1x * u1x + d1y * u1y) + (double)cfg.w_eps * det;
  const double dy = nne * (d0x * u0y - d0y * u0x + d1x * u1y - d1y * u1x);
  const double r = nne / (dx * dx + dy * dy);
  w0 = make_float2((float)((u0x * dx + u0y * dy) * r), (float)((u0y * dx - u0x * dy) * r));
  w1 = make_float2((float)((u1x * dx + u1y * dy) * r), (float)((u1y * dx - u1x * dy) * r));
}

__global__ void __launch_bounds__(kWarps * 32, AVZ_MINB_STREAM)
k512_stream_step(float* __restrict__ state, const float* __restrict__ hop_in, const float* __restrict__ noise_w,
                 const float2* __restrict__ dvec, int n_streams, int t, int t_end, float lam, AvzMvdrCfg cfg,
                 float* __restrict__ hop_out, Tables tb) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int warp = warp_id_uniform();
  float* stage = reinterpret_cast<float*>(smem_raw) + (size_t)warp * kStreamStageFloats;
  float2* sm = reinterpret_cast<float2*>(stage);   // reused once hop, history and tail are in registers
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw + (size_t)kWarps * kStreamStageFloats * sizeof(float)) + warp;
  const int s = blockIdx.x * kWarps + warp;
  if (s >= n_streams) return;
  float* st = state + (size_t)s * kStreamStateFloats;
  const float* xin_g = hop_in + (size_t)s * 256;   // [2][128]
  if ((threadIdx.x & 31) == 0) {
    mbar_init(bar, 1);
    mbar_fence_init();
    mbar_arrive_expect_tx(bar, kStreamStageFloats * sizeof(float));
    tma_load_1d(stage, xin_g, 256 * sizeof(float), bar);
    tma_load_1d(stage + 256, st, kStreamStateFloats * sizeof(float), bar);
  }
  Lane ln;
  ln.init(tb.tw);
  const int lane = ln.lane;
  float hw[16];
  load_window(hw, tb.win, ln);
  const bool frame_valid = (t >= 0) && (t < t_end);
  // this lane's noise weights and steering vectors: requested now, used after the forward transform
  float mk[9];
#pragma unroll
  for (int j = 0; j < 8; ++j) mk[j] = noise_w ? __ldg(noise_w + (size_t)s * kF + bin_lo(ln, j)) : 1.f;
  mk[8] = noise_w ? __ldg(noise_w + (size_t)s * kF + 256) : 1.f;
  __syncwarp();
  mbar_wait(bar, 0);

  const float* xin = stage;                        // [2][128]
  const float* hist = stage + 256;                 // [2][384]
  const float* tail = stage + 256 + 768;           // [384]  open output blocks t+1 .. t+3 before this frame is added
  const float* cov = stage + 256 + 1152;           // [5][288]
  float* cov_g = st + 1152;

  // frame t = [history (rows 0..11) | new hop (rows 12..15)], then slide the history
  float a[16], b[16];
#pragma unroll
  for (int r = 0; r < 12; ++r) {
    a[r] = hist[32 * r + lane];
    b[r] = hist[384 + 32 * r + lane];
  }
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    a[12 + r] = xin[32 * r + lane];
    b[12 + r] = xin[128 + 32 * r + lane];
  }
#pragma unroll
  for (int r = 0; r < 12; ++r) {
    st[32 * r + lane] = a[r + 4];
    st[384 + 32 * r + lane] = b[r + 4];
  }
  float o[16];
#pragma unroll
  for (int r = 0; r < 12; ++r) o[r] = tail[32 * r + lane];
#pragma unroll
  for (int r = 12; r < 16; ++r) o[r] = 0.f;
  __syncwarp();   // every lane has its hop, history and tail: their shared memory now serves the transposition

  if (frame_valid) {
    float2 v[16];
#pragma unroll
    for (int r = 0; r < 16; ++r) v[r] = make_float2(a[r] * hw[r], b[r] * hw[r]);
    f512::forward(v, sm, ln);
    float2 mir[8];
    f512::mirror_of_low(v, mir, ln);
    const float sc = 1.0f / ((float)kN * (float)kN);   // spectra are unscaled (x N/2) and un-halved (x 2)
    const float one_m = 1.f - lam;
    float2 S[8];
    float s_ny = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int k = bin_lo(ln, j);
      const float2 y0 = make_float2(v[j].x + mir[j].x, v[j].y - mir[j].y);   // 2 Y0 (x N/2)
      const float2 y1 = make_float2(v[j].y + mir[j].y, mir[j].x - v[j].x);   // 2 Y1
      const float m = mk[j];
      const int ci = 32 * j + lane;
      const float r00 = fmaf(lam, cov[0 * 288 + ci], one_m * m * sc * cabs2(y0));
      const float r11 = fmaf(lam, cov[1 * 288 + ci], one_m * m * sc * cabs2(y1));
      const float2 c01 = cmulc(y0, y1);
      const float rre = fmaf(lam, cov[2 * 288 + ci], one_m * m * sc * c01.x);
      const float rim = fmaf(lam, cov[3 * 288 + ci], one_m * m * sc * c01.y);
      const float nn = fmaf(lam, cov[4 * 288 + ci], one_m * m);
      cov_g[0 * 288 + ci] = r00;
      cov_g[1 * 288 + ci] = r11;
      cov_g[2 * 288 + ci] = rre;
      cov_g[3 * 288 + ci] = rim;
      cov_g[4 * 288 + ci] = nn;
      float2 w0 = make_float2(0.f, 0.f), w1 = make_float2(0.f, 0.f);
      if (k < cfg.hp_bins && cfg.hp_mode != AVZ_HP_NONE) {
        if (cfg.hp_mode == AVZ_HP_MIC0) w0.x = 1.f;
      } else {
        stream_weights(r00, r11, rre, rim, nn, __ldg(dvec + 2 * k), __ldg(dvec + 2 * k + 1), cfg, w0, w1);
      }
      // S = conj(w0) Y0 + conj(w1) Y1 with Y = y' / N (analysis 2/N, halved); synthesis factor 1/2 folded below
      const float2 sv = cadd(cmulc(y0, w0), cmulc(y1, w1));
      S[j] = make_float2(sv.x * (0.5f / (float)kN), sv.y * (0.5f / (float)kN));
    }
    // Nyquist bin (meaningful on lane 0): Y0 = Re hi[0], Y1 = Im hi[0] (x N/2, not doubled)
    {
      const float scn = 4.0f / ((float)kN * (float)kN);
      const float m = mk[8];
      const int ci = 256 + lane;    // row 8 of the cov block: only lane 0's entry is meaningful
      const float r00 = fmaf(lam, cov[0 * 288 + ci], one_m * m * scn * v[8].x * v[8].x);
      const float r11 = fmaf(lam, cov[1 * 288 + ci], one_m * m * scn * v[8].y * v[8].y);
      const float rre = fmaf(lam, cov[2 * 288 + ci], one_m * m * scn * v[8].x * v[8].y);
      const float nn = fmaf(lam, cov[4 * 288 + ci], one_m * m);
      cov_g[0 * 288 + ci] = r00;
      cov_g[1 * 288 + ci] = r11;
      cov_g[2 * 288 + ci] = rre;
      cov_g[4 * 288 + ci] = nn;
      float2 w0, w1;
      stream_weights(r00, r11, rre, 0.f, nn, __ldg(dvec + 512), __ldg(dvec + 513), cfg, w0, w1);
      // Re(conj(w0) Y0 + conj(w1) Y1) for real Y0, Y1; (2/N) analysis x 1/2 synthesis = 1/N
      s_ny = (w0.x * v[8].x + w1.x * v[8].y) * (1.0f / (float)kN);
      if (256 < cfg.hp_bins && cfg.hp_mode == AVZ_HP_ZERO) s_ny = 0.f;
    }
    float2 Z0[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) Z0[j] = make_float2(0.f, 0.f);
    float2 x[16];
    f512::hermitian_pack(S, Z0, make_float2(s_ny, 0.f), x, ln);
    f512::inverse(x, sm, ln);
#pragma unroll
    for (int r = 0; r < 16; ++r) o[r] = fmaf(hw[r], x[r].x, o[r]);
  }
  // block g = t is complete: normalise by the sum of w^2 over the frames t-3..t that exist, emit, keep the rest open
  float* yo = hop_out + (size_t)s * 128;
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    float nrm = 0.f;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int tq = t - q;
      if (tq >= 0 && tq < t_end) nrm = fmaf(hw[r + 4 * q], hw[r + 4 * q], nrm);
    }
    yo[32 * r + lane] = o[r] / (nrm > 1e-10f ? nrm : 1.0f);
  }
#pragma unroll
  for (int r = 0; r < 12; ++r) st[768 + 32 * r + lane] = o[r + 4];
}

int launch_stream_step(float* state, const float* hop_in, const float* noise_w, const float* dvec, int n_streams,
                       int t, int t_end, float lam, const AvzMvdrCfg* cfg, float* hop_out, cudaStream_t st) {
  Tables tb;
  int rc = tables_for(kN, &tb);
  if (rc) return rc;
  const size_t smem = kStreamSmem;
  AVZ_CUDA_OK(cudaFuncSetAttribute(k512_stream_step, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k512_stream_step<<<(n_streams + kWarps - 1) / kWarps, kWarps * 32, smem, st>>>(
      state, hop_in, noise_w, reinterpret_cast<const float2*>(dvec), n_streams, t, t_end, lam, *cfg, hop_out, tb);
  AVZ_LAUNCH_OK("k512_stream_step");
  return AVZ_OK;
}

}  // namespace o512
}  // namespace avz

// ------------------------------------------------------------------------------------------
// Features straight from the waveform (full_audio.../inference.py:90-94) on the fast path: STFT of the two mics,
// ln(|Y0| + 1e-7) and angle(Y0) - angle(Y1); the spectrum is never written.  A CTA covers 32 consecutive frames
// (8 per warp, sliding window), stages the two feature planes in shared memory [2][257][33] and writes every
// (feature, bin) row as one coalesced 128-byte run along time - the reference layout (2, F, T) has T contiguous.
// ------------------------------------------------------------------------------------------
namespace avz {
namespace o512 {

#ifndef AVZ_FEAT_WARPS
#define AVZ_FEAT_WARPS 8
#endif
#ifndef AVZ_MINB_FEAT
#define AVZ_MINB_FEAT 2
#endif
constexpr int kFeatWarps = AVZ_FEAT_WARPS;          // the 68 KB feature tile allows two CTAs per SM: eight warps each
constexpr int kFeatTile = 32;                       // frames per CTA
constexpr int kFeatPitch = kFeatTile + 1;           // floats per (feature, bin) row in shared memory

template <int HOP>
__global__ void __launch_bounds__(kFeatWarps * 32, AVZ_MINB_FEAT)
k512_features(const float* __restrict__ mix, int L, int T, int wrapped, float* __restrict__ X, Tables tb) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float2* sm = reinterpret_cast<float2*>(smem_raw) + (size_t)(threadIdx.x >> 5) * f512::kSmemComplex;
  float* tile = reinterpret_cast<float*>(reinterpret_cast<float2*>(smem_raw) + (size_t)kFeatWarps * f512::kSmemComplex);
  Lane ln;
  ln.init(tb.tw);
  const int lane = ln.lane, warp = warp_id_uniform();
  const int b = blockIdx.y;
  const int t0 = blockIdx.x * kFeatTile;
  const float* m0 = mix + (int64_t)b * 2 * L;
  const float* m1 = m0 + L;
  float w[16];
  load_window(w, tb.win, ln);
  constexpr int per = kFeatTile / kFeatWarps;
  const int ta = t0 + warp * per, tb_ = min(T, ta + per);
  if (ta < tb_) {
    Window2<HOP> win;
    win.load_all(m0, m1, L, ta, lane);
#pragma unroll 1
    for (int t = ta; t < tb_; ++t) {
      if (t + 1 < tb_) win.prefetch(m0, m1, L, t + 1, lane);
      float2 v[16];
      win.frame(v, w);
      f512::forward(v, sm, ln);
      float2 mir[8];
      f512::mirror_of_low(v, mir, ln);
      const int tl = t - t0;
      const float sc = 1.0f / (float)kN;   // un-halved, unscaled spectra -> true values
#pragma unroll
      for (int j = 0; j < 9; ++j) {
        float2 y0, y1;
        int k;
        if (j < 8) {
          k = bin_lo(ln, j);
          y0 = make_float2(sc * (v[j].x + mir[j].x), sc * (v[j].y - mir[j].y));
          y1 = make_float2(sc * (v[j].y + mir[j].y), sc * (mir[j].x - v[j].x));
        } else {  // Nyquist, lane 0: real spectra (imaginary parts +0 as rfft gives)
          k = 256;
          y0 = make_float2(2.f * sc * v[8].x, 0.f);
          y1 = make_float2(2.f * sc * v[8].y, 0.f);
        }
        if (j < 8 || lane == 0) {
          float lm, ipd;
          feature_values(y0, y1, lm, ipd);
          if (wrapped) {
            const float two_pi = 6.28318530717958647692f;
            ipd = ipd - two_pi * rintf(ipd / two_pi);
          }
          tile[(0 * kF + k) * kFeatPitch + tl] = lm;
          tile[(1 * kF + k) * kFeatPitch + tl] = ipd;
        }
      }
      if (t + 1 < tb_) win.advance();
    }
  }
  __syncthreads();
  const int nt = min(kFeatTile, T - t0);
  // rows (feature c, bin k): 32 consecutive frames each; one warp writes one row per iteration
  for (int row = warp; row < 2 * kF; row += kFeatWarps) {
    const int c = row / kF, k = row - c * kF;
    if (lane < nt) X[(((int64_t)b * 2 + c) * kF + k) * T + t0 + lane] = tile[row * kFeatPitch + lane];
  }
}

template <int HOP>
int launch_features(const float* mix, int B, int64_t L, int wrapped, float* X, cudaStream_t st) {
  Tables tb;
  int rc = tables_for(kN, &tb);
  if (rc) return rc;
  if (L >= (1ll << 30)) return set_error(AVZ_EINVAL, "L=%lld too long for the 512-point fast path", (long long)L);
  const int T = (int)avz_num_frames(L, kN, HOP);
  const size_t smem = (size_t)kFeatWarps * f512::kSmemComplex * sizeof(float2) + (size_t)2 * kF * kFeatPitch * sizeof(float);
  AVZ_CUDA_OK(cudaFuncSetAttribute(k512_features<HOP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid((T + kFeatTile - 1) / kFeatTile, B);
  k512_features<HOP><<<grid, kFeatWarps * 32, smem, st>>>(mix, (int)L, T, wrapped, X, tb);
  AVZ_LAUNCH_OK("k512_features");
  return AVZ_OK;
}
template int launch_features<128>(const float*, int, int64_t, int, float*, cudaStream_t);
template int launch_features<256>(const float*, int, int64_t, int, float*, cudaStream_t);

}  // namespace o512
}  // namespace avz

// Register-resident 512-point complex FFT, one warp per transform, 16 complex values per lane.
//
//   512 = 16 x 32:  radix-16 in registers  ->  twiddle  ->  ONE transposition through shared memory
//                   ->  radix-16 in registers  ->  radix-2 across lane pairs (8 shuffles of complex values)
//
// Data layouts (validated lane by lane in tools/fft512_model.py):
//   time      lane L (0..31) holds x[32 r + L], r = 0..15                       ("natural" layout)
//   spectrum  lane l = k1 + 16 h holds  lo[j] = X[k1 + 16 j + 128 h]  (bins < 256)
//                                       hi[j] = X[k1 + 16 j + 128 h + 256],  j = 0..7
//   The bin 512 - k of lo[j] lives in lane mirror(l) = ((16 - k1) & 15) + 16 (1 - h) as hi[7 - j]
//   (lanes with k1 == 0: as hi[8 - j]; their j == 0 entries are self-paired) - one shuffle per j gives every
//   lane the conjugate-symmetric partner of each of its low bins, which is all the real-signal algebra
//   (two real frames per complex transform) needs.
//
// Shared memory per warp: 16 rows x 42 complex (5376 B); row k1, column (n2 & 1) * 24 + (n2 >> 1) holds
// element (k1, n2).  8-byte stores by lane n2 and 16-byte loads of 16 consecutive columns by lane (k1, h) are
// both bank-conflict free.  Only __syncwarp() is needed: no block-level barrier on the FFT path.
#pragma once
#include "avz_common.cuh"

namespace avz {
namespace f512 {

constexpr int kN = 512;
constexpr int kRow = 42;                  // complex elements per shared-memory row
constexpr int kSmemComplex = 16 * kRow;   // per warp

template <int M16>
struct W16 {  // exp(-2 pi i M16 / 16), compile-time
  static constexpr float c() {
    constexpr float t[16] = {1.f, 0.92387953251128674f, 0.70710678118654752f, 0.38268343236508977f,
                             0.f, -0.38268343236508977f, -0.70710678118654752f, -0.92387953251128674f,
                             -1.f, -0.92387953251128674f, -0.70710678118654752f, -0.38268343236508977f,
                             0.f, 0.38268343236508977f, 0.70710678118654752f, 0.92387953251128674f};
    return t[M16 & 15];
  }
  static constexpr float s() { return W16<(M16 + 12) & 15>::c(); }  // sin(x) = cos(x - pi/2)
};

template <int M32>
struct W32 {  // exp(-2 pi i M32 / 32) for M32 = 0..7
  static constexpr float c() {
    constexpr float t[8] = {1.f, 0.98078528040323043f, 0.92387953251128674f, 0.83146961230254524f,
                            0.70710678118654752f, 0.55557023301960218f, 0.38268343236508977f, 0.19509032201612825f};
    return t[M32];
  }
  static constexpr float s() {
    constexpr float t[8] = {0.f, 0.19509032201612825f, 0.38268343236508977f, 0.55557023301960218f,
                            0.70710678118654752f, 0.83146961230254524f, 0.92387953251128674f, 0.98078528040323043f};
    return t[M32];
  }
};

// v * exp(-+ 2 pi i M / 16)   (forward: minus)
template <int M, bool INV>
__device__ __forceinline__ float2 mul_w16(float2 v) {
  constexpr int m = M & 15;
  if (m == 0) return v;
  if (m == 4) return INV ? make_float2(-v.y, v.x) : make_float2(v.y, -v.x);
  if (m == 8) return make_float2(-v.x, -v.y);
  if (m == 12) return INV ? make_float2(v.y, -v.x) : make_float2(-v.y, v.x);
  constexpr float c = W16<m>::c();
  constexpr float s = INV ? -W16<m>::s() : W16<m>::s();  // w = c - i s (forward)
  // (x + i y)(c - i s) = (x, y) c + (y, -x) s : one packed multiply + one packed fma (swap and sign are operand modifiers)
  return up2(fma2(pk2(v.y, -v.x), pk2(s, s), mul2(pk2(v), pk2(c, c))));
}

template <bool INV>
__device__ __forceinline__ void radix4(float2& x0, float2& x1, float2& x2, float2& x3) {
  const float2 t0 = cadd(x0, x2), t1 = csub(x0, x2), t2 = cadd(x1, x3), t3 = csub(x1, x3);
  x0 = cadd(t0, t2);
  x2 = csub(t0, t2);
  // t1 -+ i t3 (forward) / t1 +- i t3 (inverse): the rotation by +-i is the half swap + sign of the packed operand
  const u64 p1 = pk2(t1);
  x1 = up2(add2(p1, INV ? pk2(-t3.y, t3.x) : pk2(t3.y, -t3.x)));
  x3 = up2(add2(p1, INV ? pk2(t3.y, -t3.x) : pk2(-t3.y, t3.x)));
}

// 16-point DFT in registers, natural order in and out (4 x 4 decomposition).
template <bool INV>
__device__ __forceinline__ void fft16(float2 (&v)[16]) {
  // step 1: radix-4 over n1 for each n2 (elements 4 n1 + n2); results a[n2][k1] stay at v[4 k1 + n2]
#pragma unroll
  for (int n2 = 0; n2 < 4; ++n2) radix4<INV>(v[n2], v[4 + n2], v[8 + n2], v[12 + n2]);
  // twiddle a[n2][k1] *= W16^(n2 k1)
  v[4 * 1 + 1] = mul_w16<1, INV>(v[4 * 1 + 1]);
  v[4 * 1 + 2] = mul_w16<2, INV>(v[4 * 1 + 2]);
  v[4 * 1 + 3] = mul_w16<3, INV>(v[4 * 1 + 3]);
  v[4 * 2 + 1] = mul_w16<2, INV>(v[4 * 2 + 1]);
  v[4 * 2 + 2] = mul_w16<4, INV>(v[4 * 2 + 2]);
  v[4 * 2 + 3] = mul_w16<6, INV>(v[4 * 2 + 3]);
  v[4 * 3 + 1] = mul_w16<3, INV>(v[4 * 3 + 1]);
  v[4 * 3 + 2] = mul_w16<6, INV>(v[4 * 3 + 2]);
  v[4 * 3 + 3] = mul_w16<9, INV>(v[4 * 3 + 3]);
  // step 2: radix-4 over n2 for each k1 (elements 4 k1 + n2) -> X[k1 + 4 k2]
  float2 o[16];
#pragma unroll
  for (int k1 = 0; k1 < 4; ++k1) {
    float2 a = v[4 * k1 + 0], b = v[4 * k1 + 1], c = v[4 * k1 + 2], d = v[4 * k1 + 3];
    radix4<INV>(a, b, c, d);
    o[k1] = a;
    o[k1 + 4] = b;
    o[k1 + 8] = c;
    o[k1 + 12] = d;
  }
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = o[i];
}

// Per-lane constants of the transform.
//
// The transposition twiddle of element k (0..15) of lane L is W512^(L k).  With k = k1 + 4 k2 it factors into
// ta[k1] * tb[k2], ta[k1] = W512^(L k1), tb[k2] = W512^(4 L k2): 6 complex registers instead of 15, at the price
// of 9 extra complex multiplies per transform (registers, not flops, limit occupancy here).
// Lanes L = 3 (mod 4) must also negate their 16 values (this rotates the second-stage outputs of the odd lanes
// by 8 so that the radix-2 exchange needs no register selects); callers fold that `sign` into the analysis /
// synthesis window they multiply with anyway.
struct Lane {
  float2 ta[4], tb[4];   // [0] unused (= 1)
  float2 tf[16];         // all 15 products ta[k1] * tb[k2] (init_full(); only kernels with registers to spare use them)
  float rc[8], rs[8];    // radix-2 twiddles W32^j * (-i)^h as per-lane constants (init_full(); [0] unused)
  float sign;            // -1 on lanes L = 3 (mod 4), else +1: multiply the time-domain side by it
  int lane, k1, h, mirror;
  bool k0;               // k1 == 0: the mirrored bins sit one slot further (see header comment)
  float rot;             // h ? -1 : +1
  int wr_off;            // shared-memory column this lane writes in the forward transposition
  int rd_off;            // first element this lane reads

  __device__ __forceinline__ void init(const float2* __restrict__ tw512 /* exp(-2 pi i k/512) */) {
    lane = threadIdx.x & 31;
    k1 = lane & 15;
    h = lane >> 4;
    mirror = ((16 - k1) & 15) + 16 * (1 - h);
    k0 = (k1 == 0);
    rot = h ? -1.f : 1.f;
    sign = ((lane & 3) == 3) ? -1.f : 1.f;
#pragma unroll
    for (int i = 1; i < 4; ++i) {
      ta[i] = tw512[(lane * i) & 511];
      tb[i] = tw512[(lane * 4 * i) & 511];
    }
    ta[0] = tb[0] = make_float2(1.f, 0.f);
    wr_off = (lane & 1) * 24 + (lane >> 1);
    rd_off = k1 * kRow + 24 * h;
  }

  // the unfactored twiddles W512^(lane k), k = 1..15: 18 more registers, 9 fewer complex multiplies per transform
  __device__ __forceinline__ void init_full(const float2* __restrict__ tw512) {
#pragma unroll
    for (int k = 1; k < 16; ++k) tf[k] = tw512[(lane * k) & 511];
    tf[0] = make_float2(1.f, 0.f);
#pragma unroll
    for (int j = 0; j < 8; ++j) {   // W32^j = tw512[16 j] = (c, -s);  h = 1: (c', s') = (-s, c)
      const float2 w = tw512[16 * j];
      rc[j] = h ? w.y : w.x;
      rs[j] = h ? w.x : -w.y;
    }
  }
};

// Second half of the forward transform (after the transposition twiddles have been applied).  LANEC: the radix-2
// twiddles come from Lane::rc/rs (14 registers) instead of immediates + 2 selects + 1 multiply per j.
template <bool LANEC>
__device__ __forceinline__ void forward_tail(float2 (&v)[16], float2* __restrict__ sm, const Lane& ln);

// Forward transform with the 15 unfactored twiddles (Lane::init_full).
__device__ __forceinline__ void forward_full(float2 (&v)[16], float2* __restrict__ sm, const Lane& ln) {
  fft16<false>(v);
#pragma unroll
  for (int k = 1; k < 16; ++k) v[k] = cmul(v[k], ln.tf[k]);
  forward_tail<true>(v, sm, ln);
}

// Forward transform.  In: v[r] = sign * x[32 r + lane].  Out: v[j] = lo[j], v[8 + j] = hi[j] (spectrum layout).
__device__ __forceinline__ void forward(float2 (&v)[16], float2* __restrict__ sm, const Lane& ln) {
  fft16<false>(v);
#pragma unroll
  for (int a = 1; a < 4; ++a)
#pragma unroll
    for (int c = 0; c < 4; ++c) v[a + 4 * c] = cmul(v[a + 4 * c], ln.ta[a]);
#pragma unroll
  for (int c = 1; c < 4; ++c)
#pragma unroll
    for (int a = 0; a < 4; ++a) v[a + 4 * c] = cmul(v[a + 4 * c], ln.tb[c]);
  forward_tail<false>(v, sm, ln);
}

template <bool LANEC>
__device__ __forceinline__ void forward_tail(float2 (&v)[16], float2* __restrict__ sm, const Lane& ln) {
  float2* wp = sm + ln.wr_off;
#pragma unroll
  for (int i = 0; i < 16; ++i) wp[i * kRow] = v[i];
  __syncwarp();
  const float4* rp = reinterpret_cast<const float4*>(sm + ln.rd_off);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float4 q = rp[i];
    v[2 * i] = make_float2(q.x, q.y);
    v[2 * i + 1] = make_float2(q.z, q.w);
  }
  __syncwarp();
  fft16<false>(v);  // h = 1 lanes: v[q'] = C_1[(q' + 8) % 16] (inputs were sign-modulated by the writers)
  const bool hh = ln.h != 0;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float rx = __shfl_xor_sync(kFull, v[8 + j].x, 16);
    const float ry = __shfl_xor_sync(kFull, v[8 + j].y, 16);
    const float2 E = hh ? make_float2(rx, ry) : v[j];
    const float2 O = hh ? v[j] : make_float2(rx, ry);
    if (LANEC && j > 0) {   // t = O * W32^j * (-i)^h with this lane's constants: (O.x, O.y) rc + (O.y, -O.x) rs
      const float2 t = up2(fma2(pk2(O.y, -O.x), pk2(ln.rs[j], ln.rs[j]), mul2(pk2(O), pk2(ln.rc[j], ln.rc[j]))));
      v[j] = cadd(E, t);
      v[8 + j] = csub(E, t);
      continue;
    }
    // t = O * W32^j ; for h = 1 the twiddle is W32^(j+8) = -i W32^j
    float2 t;
    if (j == 0) {
      t = O;
    } else {
      const float c = (j == 1) ? W32<1>::c() : (j == 2) ? W32<2>::c() : (j == 3) ? W32<3>::c() : (j == 4) ? W32<4>::c()
                    : (j == 5) ? W32<5>::c() : (j == 6) ? W32<6>::c() : W32<7>::c();
      const float s = (j == 1) ? W32<1>::s() : (j == 2) ? W32<2>::s() : (j == 3) ? W32<3>::s() : (j == 4) ? W32<4>::s()
                    : (j == 5) ? W32<5>::s() : (j == 6) ? W32<6>::s() : W32<7>::s();
      t = up2(fma2(pk2(O.y, -O.x), pk2(s, s), mul2(pk2(O), pk2(c, c))));
    }
    const float p = hh ? t.y : t.x;            // (-i)^h t = (p, q)
    const float q = (hh ? t.x : t.y) * ln.rot;
    v[j] = cadd(E, make_float2(p, q));
    v[8 + j] = csub(E, make_float2(p, q));
  }
}

// Unnormalised inverse transform.  In: v[j] = lo[j], v[8 + j] = hi[j].  Out: v[r] = sign * 512 * x[32 r + lane].
// FULL: the 15 unfactored twiddles of Lane::init_full instead of the 6 factored ones.
template <bool FULL = false>
__device__ __forceinline__ void inverse(float2 (&v)[16], float2* __restrict__ sm, const Lane& ln) {
  const bool hh = ln.h != 0;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float2 Ep = cadd(v[j], v[8 + j]);
    const float2 d = csub(v[j], v[8 + j]);
    // Op = d * conj(W32^(j + 8 h)) = d * conj(W32^j) * (+i)^h
    float2 t;
    if (j == 0) {
      t = d;
    } else {
      const float c = (j == 1) ? W32<1>::c() : (j == 2) ? W32<2>::c() : (j == 3) ? W32<3>::c() : (j == 4) ? W32<4>::c()
                    : (j == 5) ? W32<5>::c() : (j == 6) ? W32<6>::c() : W32<7>::c();
      const float s = (j == 1) ? W32<1>::s() : (j == 2) ? W32<2>::s() : (j == 3) ? W32<3>::s() : (j == 4) ? W32<4>::s()
                    : (j == 5) ? W32<5>::s() : (j == 6) ? W32<6>::s() : W32<7>::s();
      // (x + i y)(c + i s) = (x, y) c + (-y, x) s
      t = up2(fma2(pk2(-d.y, d.x), pk2(s, s), mul2(pk2(d), pk2(c, c))));
    }
    // (+i)^h t: h = 1 -> (-t.y, t.x)
    const float2 Op = make_float2((hh ? t.y : t.x) * ln.rot, hh ? t.x : t.y);
    const float2 send = hh ? Ep : Op;
    const float2 keep = hh ? Op : Ep;
    v[j] = keep;
    v[8 + j].x = __shfl_xor_sync(kFull, send.x, 16);
    v[8 + j].y = __shfl_xor_sync(kFull, send.y, 16);
  }
  fft16<true>(v);
  float4* wp = reinterpret_cast<float4*>(sm + ln.rd_off);
#pragma unroll
  for (int i = 0; i < 8; ++i) wp[i] = make_float4(v[2 * i].x, v[2 * i].y, v[2 * i + 1].x, v[2 * i + 1].y);
  __syncwarp();
  const float2* rp = sm + ln.wr_off;
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = rp[i * kRow];
  __syncwarp();
  if (FULL) {
#pragma unroll
    for (int k = 1; k < 16; ++k) v[k] = cmulc(v[k], ln.tf[k]);
  } else {
#pragma unroll
    for (int a = 1; a < 4; ++a)
#pragma unroll
      for (int c = 0; c < 4; ++c) v[a + 4 * c] = cmulc(v[a + 4 * c], ln.ta[a]);
#pragma unroll
    for (int c = 1; c < 4; ++c)
#pragma unroll
      for (int a = 0; a < 4; ++a) v[a + 4 * c] = cmulc(v[a + 4 * c], ln.tb[c]);
  }
  fft16<true>(v);
}

// mir[j] = X[(512 - k) mod 512] for k the bin of lo[j] = v[j]  (hi = v[8..15]).
__device__ __forceinline__ void mirror_of_low(const float2 (&v)[16], float2 (&mir)[8], const Lane& ln) {
  float2 r[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    r[j].x = __shfl_sync(kFull, v[8 + 7 - j].x, ln.mirror);
    r[j].y = __shfl_sync(kFull, v[8 + 7 - j].y, ln.mirror);
  }
#pragma unroll
  for (int j = 1; j < 8; ++j) mir[j] = ln.k0 ? r[j - 1] : r[j];
  const float2 self0 = ln.h ? v[8] : v[0];
  mir[0] = ln.k0 ? self0 : r[0];
}

// Build the spectrum of (xa + i xb) for two real frames from their one-sided spectra held in the low-bin
// layout (Sa[j], Sb[j] at this lane's lo bins; `ny` = (Re Sa[256], Re Sb[256]), used by lane 0 only).
__device__ __forceinline__ void hermitian_pack(const float2 (&Sa)[8], const float2 (&Sb)[8], float2 ny, float2 (&v)[16],
                                               const Lane& ln) {
  float2 gm[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    float2 a = Sa[j], b = Sb[j];
    if (j == 0 && ln.lane == 0) {  // DC: the c2r transform ignores the imaginary part
      a.y = 0.f;
      b.y = 0.f;
    }
    v[j] = up2(add2(pk2(a), pk2(-b.y, b.x)));          // Sa + i Sb
    gm[j] = up2(add2(pk2(a.x, -a.y), pk2(b.y, b.x)));  // conj(Sa) + i conj(Sb): value at the mirrored bin
  }
  float2 r[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    r[j].x = __shfl_sync(kFull, gm[7 - j].x, ln.mirror);
    r[j].y = __shfl_sync(kFull, gm[7 - j].y, ln.mirror);
  }
#pragma unroll
  for (int j = 1; j < 8; ++j) v[8 + j] = ln.k0 ? r[j - 1] : r[j];
  const float2 self0 = ln.h ? gm[0] : ny;   // lane 16: mirror of bin 128 is its own hi[0]; lane 0: Nyquist
  v[8] = ln.k0 ? self0 : r[0];
}

}  // namespace f512
}  // namespace avz

"""numpy model of the cluster-resident far-field mixer (csrc/avz_mixer_cluster.cu): one in-place mixed-radix
decimation-in-frequency transform of length L = 4 * M whose first radix-4 stage runs across the four CTAs of a plane
(each CTA owns a contiguous quarter of the signal), the digit-reversed spectrum order it leaves, the mirror-position
table the Hermitian combine needs, and the inverse (decimation in time, stages backwards).  Checked against numpy's FFT
and the float64 oracle mixer.  Design aid, CPU only:  python tools/mixer_cluster_model.py
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from oracle import mvdr_oracle as O  # noqa: E402

MAX_LOCAL = 25600          # complex elements of one CTA's quarter that fit its shared memory


def plan(L):
    """Radices of the transform: 4 (across the CTAs of a plane), then the local factors of M = L / 4, powers of two
    first (16, 8, 4, 2), then 3, 5, 7.  None when L has no such split."""
    if L % 4 or L // 4 > MAX_LOCAL or L < 8:
        return None
    M, rad = L // 4, [4]
    while M % 16 == 0:
        rad.append(16)
        M //= 16
    for r in (8, 4, 2):
        if M % r == 0:
            rad.append(r)
            M //= r
    for r in (3, 5, 7):
        while M % r == 0:
            rad.append(r)
            M //= r
    return rad if M == 1 else None


def strides(L, rad):
    """q[s] = distance between the inputs of a stage-s butterfly; the stage works on blocks of R[s] * q[s]."""
    q, n = [], L
    for r in rad:
        n //= r
        q.append(n)
    return q


def dif(x, rad):
    """In-place DIF: bin k = d0 + R0 (d1 + R1 (d2 + ...)) ends at position d0 q0 + d1 q1 + ..."""
    L = len(x)
    x = x.astype(complex).copy()
    for r, q in zip(rad, strides(L, rad)):
        nb = r * q
        blk = x.reshape(L // nb, r, q)                     # [block][t][pos]
        u = np.arange(r)
        y = np.einsum("ut,btp->bup", np.exp(-2j * np.pi * np.outer(u, u) / r), blk)
        y *= np.exp(-2j * np.pi * np.arange(q)[None, None, :] * u[None, :, None] / nb)
        x = y.reshape(L)
    return x


def dit_inverse(x, rad):
    """Exact reverse of dif() up to the factor L: conj twiddle on the way in, conj butterfly, stages backwards."""
    L = len(x)
    x = x.astype(complex).copy()
    for r, q in reversed(list(zip(rad, strides(L, rad)))):
        nb = r * q
        blk = x.reshape(L // nb, r, q)
        u = np.arange(r)
        z = blk * np.exp(+2j * np.pi * np.arange(q)[None, None, :] * u[None, :, None] / nb)
        y = np.einsum("tu,bup->btp", np.exp(+2j * np.pi * np.outer(u, u) / r), z)
        x = y.reshape(L)
    return x


def bin_of_position(L, rad):
    q = strides(L, rad)
    pos = np.arange(L)
    k, mult = np.zeros(L, dtype=np.int64), 1
    for r, qs in zip(rad, q):
        d = (pos // qs) % r
        k += d * mult
        mult *= r
    return k


def model_mix(src, delays, fs):
    S, L = src.shape
    rad = plan(L)
    assert rad is not None and S <= 4
    kpos = bin_of_position(L, rad)
    pos_of_bin = np.empty(L, dtype=np.int64)
    pos_of_bin[kpos] = np.arange(L)
    mirror = pos_of_bin[(L - kpos) % L]                     # position of bin L - k
    Z = [dif(src[2 * p] + 1j * (src[2 * p + 1] if 2 * p + 1 < S else 0.0), rad) if 2 * p < S else np.zeros(L, complex)
         for p in range(2)]
    c1 = [delays[s][0] * fs / L for s in range(S)]
    c2 = [delays[s][1] * fs / L for s in range(S)]
    Q = np.zeros((2, L), complex)
    for o in range(L):
        k, om = int(kpos[o]), int(mirror[o])
        km = (L - k) % L
        if k > km:
            continue
        m1 = m2 = tg = 0j
        for p in range(2):
            zk, zm = Z[p][o], Z[p][om]
            a = 0.5 * (zk + zm.conjugate())
            bv = -0.5j * (zk - zm.conjugate())
            for h in range(2):
                s = 2 * p + h
                if s < S:
                    v = bv if h else a
                    d1 = v * np.exp(-2j * np.pi * k * c1[s])
                    d2 = v * np.exp(-2j * np.pi * k * c2[s])
                    m1 += d1
                    m2 += d2
                    if s == 0:
                        tg = d1
        it = m1 - tg
        if k == km:
            m1, m2, tg, it = m1.real + 0j, m2.real + 0j, tg.real + 0j, it.real + 0j
        Q[0][o], Q[1][o] = m1 + 1j * m2, tg + 1j * it
        if om != o:
            Q[0][om], Q[1][om] = m1.conjugate() + 1j * m2.conjugate(), tg.conjugate() + 1j * it.conjugate()
    outs = [dit_inverse(Q[q], rad) / L for q in range(2)]
    mix = np.stack([outs[0].real, outs[0].imag])
    norm = np.max(np.abs(mix)) + 1e-9
    return mix / norm, outs[1].real / norm, outs[1].imag / norm


def main():
    rng = np.random.default_rng(0)
    fs = 16000.0
    for L in (64000, 80000, 16000, 4096, 96, 512, 1000, 375, 32000, 40, 8, 12):
        print(L, plan(L))
    for L, S in ((2000, 3), (1000, 4), (96, 1), (640, 2), (512, 3), (8, 2), (840, 4), (64, 4)):
        rad = plan(L)
        src = rng.standard_normal((S, L))
        angles = [90.0, 40.0, 130.0, 65.0][:S]
        delays = [O.far_field_delays(a, 0.04, 343.0) for a in angles]
        z = src[0] + 1j * src[-1]
        X = dif(z, rad)
        ref = np.fft.fft(z)
        e_fft = np.abs(X - ref[bin_of_position(L, rad)]).max()
        e_inv = np.abs(dit_inverse(X, rad) / L - z).max()
        mix, tgt, itf = model_mix(src, delays, fs)
        rmix, rtgt, ritf = O.mix_far_field(list(src), angles, 0.04, 343.0, fs)
        e = max(np.abs(mix - rmix).max(), np.abs(tgt - rtgt).max(), np.abs(itf - ritf).max())
        print(f"L={L} S={S} radices {rad}  fft err {e_fft:.2e}  round trip {e_inv:.2e}  mixer err {e:.2e}")
        assert e_fft < 1e-9 and e_inv < 1e-12 and e < 1e-10
    # which CTA of a plane holds the mirror of a bin: quarter u holds the bins k = u (mod 4)
    L = 64000
    rad = plan(L)
    kpos = bin_of_position(L, rad)
    M = L // 4
    assert all(np.all(kpos[u * M:(u + 1) * M] % 4 == u) for u in range(4))


if __name__ == "__main__":
    main()

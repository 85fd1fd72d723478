"""numpy model of the cluster-resident far-field mixer (csrc/avz_mixer_cluster.cuh, `k_mix_cluster`), as it ships:

* the two complex planes of an utterance are split over 2 x 4 CTAs, CTA (p, a) holding the DECIMATED sequence
  x[4 m + a], m < M = L / 4, of plane p;
* each CTA transforms its own M values in place (mixed-radix decimation in frequency, radices 16 / 8 / 4 / 2 / 3 / 7 /
  5, a trailing 5 x 5 fused into one radix-25 stage whose block is kept in natural order of its digit), which leaves
  local bin k' at a digit-reversed position;
* ONE exchange step does, per pair of local bins (k', M - k'): the cross-CTA radix-4 butterfly
  X[k' + M u] = sum_a W4^{a u} W_L^{a k'} Y_a[k'], the Hermitian unpack / phase ramps / re-pack of the four mirror pairs
  among the eight bins k' + M u, (M - k') + M u, the inverse cross butterfly, and writes the 16 values back in place;
* the inverse local transforms run the stages backwards (decimation in time) and land in natural order;
* global memory <-> decimated sequences: CTA a of a plane moves the contiguous quarter m in [a M / 4, (a + 1) M / 4) of
  all four sequences (float4 = x[4 m .. 4 m + 3]) and sample j travels to / from CTA j's position m.

Checked against numpy's FFT and the float64 oracle mixer; also prints how contiguous the positions of a task list are
(what the natural-order radix-25 block buys: DSMEM moves 32-byte sectors).  Design aid, CPU only:
    python tools/mixer_cluster_model.py
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from oracle import mvdr_oracle as O  # noqa: E402

MAX_LOCAL = 25600          # complex elements of one CTA's sequence that fit its shared memory


def plan(L):
    """Local radices of M = L / 4 (cluster_radices): 16s, then one of 8 / 4 / 2 each, then 3s, 7s, 5s; a trailing 5, 5
    becomes one radix-25 stage.  None when L has no such split."""
    if L % 4 or L // 4 > MAX_LOCAL or L < 8:
        return None
    m, rad = L // 4, []
    while m % 16 == 0:
        rad.append(16)
        m //= 16
    for r in (8, 4, 2):
        if m % r == 0:
            rad.append(r)
            m //= r
    for r in (3, 7, 5):
        while m % r == 0:
            rad.append(r)
            m //= r
    if m != 1 or not rad:
        return None
    if len(rad) >= 2 and rad[-1] == 5 and rad[-2] == 5:
        rad[-2:] = [25]
    return rad


def strides(M, rad):
    q, n = [], M
    for r in rad:
        n //= r
        q.append(n)
    return q


def dif_local(x, rad):
    """In-place DIF on one CTA's M values: stage s works on blocks of R q, y[u] = sum_t x[t] W_R^{ut}, times W_{Rq}^{pos u};
    local bin k' = d0 + R0 (d1 + R1 (...)) ends at position d0 q0 + d1 q1 + ...  (a radix-25 stage is an ordinary stage
    here: its block comes out in natural order of the digit, which is what the kernel's register transposition does)."""
    M = len(x)
    x = x.astype(complex).copy()
    for r, q in zip(rad, strides(M, rad)):
        nb = r * q
        blk = x.reshape(M // nb, r, q)
        u = np.arange(r)
        y = np.einsum("ut,btp->bup", np.exp(-2j * np.pi * np.outer(u, u) / r), blk)
        y *= np.exp(-2j * np.pi * np.arange(q)[None, None, :] * u[None, :, None] / nb)
        x = y.reshape(M)
    return x


def dit_local_inverse(x, rad):
    """The exact reverse: conjugate twiddle on the way in, conjugate butterfly, stages backwards; natural order out."""
    M = len(x)
    x = x.astype(complex).copy()
    for r, q in reversed(list(zip(rad, strides(M, rad)))):
        nb = r * q
        blk = x.reshape(M // nb, r, q).copy()
        u = np.arange(r)
        blk *= np.exp(2j * np.pi * np.arange(q)[None, None, :] * u[None, :, None] / nb)
        y = np.einsum("tu,bup->btp", np.exp(2j * np.pi * np.outer(u, u) / r), blk)
        x = y.reshape(M)
    return x


def bin_of_position(M, rad):
    pos = np.arange(M)
    k, mult = np.zeros(M, dtype=np.int64), 1
    for r, q in zip(rad, strides(M, rad)):
        k += ((pos // q) % r) * mult
        mult *= r
    return k


def task_list(M, rad):
    """(position of k', position of M - k', k') for k' <= M - k', sorted by the first position (cluster_plan_for)."""
    kpos = bin_of_position(M, rad)
    pos_of = np.empty(M, dtype=np.int64)
    pos_of[kpos] = np.arange(M)
    out = []
    for pos in range(M):
        k = int(kpos[pos])
        kb = (M - k) % M
        if k <= kb:
            out.append((pos, int(pos_of[kb]), k))
    return out


def combine(zk, zm, k, S, c1, c2, self_pair):
    """Hermitian unpack / ramps / re-pack of the mirror pair (k, L - k) on both planes (ClCombine)."""
    m1 = m2 = tg = 0j
    for p in range(2):
        a = 0.5 * (zk[p] + zm[p].conjugate())
        bv = -0.5j * (zk[p] - zm[p].conjugate())
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
    if self_pair:
        m1, m2, tg, it = m1.real + 0j, m2.real + 0j, tg.real + 0j, it.real + 0j
    return ([m1 + 1j * m2, tg + 1j * it], [m1.conjugate() + 1j * m2.conjugate(), tg.conjugate() + 1j * it.conjugate()])


def model_mix(src, delays, fs):
    S, L = src.shape
    rad = plan(L)
    assert rad is not None and S <= 4
    M = L // 4
    # "XIO" load: loader a handles m in [a M / 4, (a + 1) M / 4) and hands sample j of x[4 m .. 4 m + 3] to CTA j
    sm = np.zeros((2, 4, M), complex)
    for p in range(2):
        if 2 * p < S:
            z = src[2 * p] + 1j * (src[2 * p + 1] if 2 * p + 1 < S else 0.0)
            for a in range(4):
                for m in range(M * a // 4, M * (a + 1) // 4):
                    for j in range(4):
                        sm[p, j, m] = z[4 * m + j]
            for a in range(4):
                sm[p, a] = dif_local(sm[p, a], rad)
    c1 = [delays[s][0] * fs / L for s in range(S)]
    c2 = [delays[s][1] * fs / L for s in range(S)]
    kpos = bin_of_position(M, rad)
    wx = np.exp(-2j * np.pi * np.outer(np.arange(4), kpos) / L)          # W_L^{a k'(pos)}
    w4 = np.exp(-2j * np.pi * np.outer(np.arange(4), np.arange(4)) / 4)
    for jA, jB, kA in task_list(M, rad):
        two = jA != jB
        XA = np.array([w4 @ (sm[p, :, jA] * wx[:, jA]) for p in range(2)])    # XA[p][u] = bin kA + M u
        XB = np.array([w4 @ (sm[p, :, jB] * wx[:, jB]) for p in range(2)])    # bin kB + M u, kB = M - kA

        def pair(X, u, Y, v, kk, self_pair):
            zk, zm = combine([X[0][u], X[1][u]], [Y[0][v], Y[1][v]], kk, S, c1, c2, self_pair)
            if not self_pair:
                Y[0][v], Y[1][v] = zm
            X[0][u], X[1][u] = zk

        if two:
            kB = M - kA
            pair(XA, 0, XB, 3, kA, False)
            pair(XA, 1, XB, 2, kA + M, False)
            pair(XB, 0, XA, 3, kB, False)
            pair(XB, 1, XA, 2, kB + M, False)
        elif kA == 0:
            pair(XA, 0, XA, 0, 0, True)
            pair(XA, 2, XA, 2, 2 * M, True)
            pair(XA, 1, XA, 3, M, False)
        else:
            pair(XA, 0, XA, 3, kA, False)
            pair(XA, 1, XA, 2, kA + M, False)
        for p in range(2):
            sm[p, :, jA] = (w4.conj() @ XA[p]) * wx[:, jA].conj() / L
            if two:
                sm[p, :, jB] = (w4.conj() @ XB[p]) * wx[:, jB].conj() / L
    outs = np.zeros((2, L), complex)
    for p in range(2):
        for a in range(4):
            outs[p, a::4] = dit_local_inverse(sm[p, a], rad)
    mix = np.stack([outs[0].real, outs[0].imag])
    norm = np.max(np.abs(mix)) + 1e-9
    return mix / norm, outs[1].real / norm, outs[1].imag / norm


def sector_report(L):
    """32-byte sectors (4 positions of 8 bytes) a warp of 32 consecutive tasks touches in one peer, per payload sector."""
    rad = plan(L)
    M = L // 4
    tasks = task_list(M, rad)
    touched = payload = 0
    for w0 in range(0, len(tasks), 32):
        w = tasks[w0:w0 + 32]
        for col in (0, 1):
            pos = {t[col] for t in w}
            touched += len({p // 4 for p in pos})
            payload += len(pos) / 4.0
    return touched / payload


def main():
    rng = np.random.default_rng(0)
    fs = 16000.0
    for L in (64000, 80000, 16000, 4096, 96, 512, 1000, 32000, 40, 8, 12):
        print(L, plan(L))
    for L, S in ((2000, 3), (1000, 4), (96, 1), (640, 2), (512, 3), (8, 2), (840, 4), (64, 4), (400, 4)):
        rad = plan(L)
        M = L // 4
        src = rng.standard_normal((S, L))
        angles = [90.0, 40.0, 130.0, 65.0][:S]
        delays = [O.far_field_delays(a, 0.04, 343.0) for a in angles]
        z = rng.standard_normal(M) + 1j * rng.standard_normal(M)
        X = dif_local(z, rad)
        e_fft = np.abs(X - np.fft.fft(z)[bin_of_position(M, rad)]).max()
        e_inv = np.abs(dit_local_inverse(X, rad) / M - z).max()
        mix, tgt, itf = model_mix(src, delays, fs)
        rmix, rtgt, ritf = O.mix_far_field(list(src), angles, 0.04, 343.0, fs)
        e = max(np.abs(mix - rmix).max(), np.abs(tgt - rtgt).max(), np.abs(itf - ritf).max())
        print(f"L={L} S={S} local radices {rad}  fft err {e_fft:.2e}  round trip {e_inv:.2e}  mixer err {e:.2e}")
        assert e_fft < 1e-9 and e_inv < 1e-12 and e < 1e-10
    print("sectors touched per payload sector in the exchange step, L = 64000: %.2f" % sector_report(64000))


if __name__ == "__main__":
    main()

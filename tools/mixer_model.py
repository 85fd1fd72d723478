"""numpy model of csrc/avz_mixer.cu's index algebra (two-factor FFT, radix-4 DIF/DIT digit reversal, packed Hermitian
combine), checked against the float64 oracle mixer.  Design aid, CPU only:  python tools/mixer_model.py
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from oracle import mvdr_oracle as O  # noqa: E402


def split_length(L, max_n1=512):
    n1, lg = 1, 0
    while n1 < max_n1 and L % (2 * n1) == 0:
        n1 *= 2
        lg += 1
    return n1, lg, L // n1


def stages(N1):
    """Radix-4 stages (quarter sizes N1/4, N1/16, ...) plus one radix-2 stage when log2 N1 is odd."""
    st, rem = [], N1
    while rem >= 4:
        st.append(("r4", rem // 4))
        rem //= 4
    if rem == 2:
        st.append(("r2", 1))
    return st


def pos_to_bin(pos, N1):
    k, mult, rem, p = 0, 1, N1, pos
    while rem >= 4:
        q = rem // 4
        d = p // q
        p -= d * q
        k += d * mult
        mult *= 4
        rem = q
    if rem == 2:
        k += p * mult
    return k


def bin_to_pos(k, N1):
    pos, rem = 0, N1
    while rem >= 4:
        q = rem // 4
        pos += (k & 3) * q
        k >>= 2
        rem = q
    if rem == 2:
        pos += k & 1
    return pos


def cols_fwd(z, W, N1, lg, N2):
    """z (N,) complex natural order -> A [k1][n2] after the in-place DIF over n1 and the W_L^{n2 k1} twiddle."""
    sm = z.reshape(N1, N2).copy()  # [n1][n2]
    for kind, q in stages(N1):
        if kind == "r4":
            ts = N2 * (N1 // (4 * q))
            for g in range(0, N1, 4 * q):
                for pos in range(q):
                    i = g + pos
                    a0, a1, a2, a3 = sm[i].copy(), sm[i + q].copy(), sm[i + 2 * q].copy(), sm[i + 3 * q].copy()
                    t0, t1, t2, t3 = a0 + a2, a0 - a2, a1 + a3, (a1 - a3) * (-1j)
                    sm[i] = t0 + t2
                    sm[i + q] = (t1 + t3) * W[pos * ts]
                    sm[i + 2 * q] = (t0 - t2) * W[2 * pos * ts]
                    sm[i + 3 * q] = (t1 - t3) * W[3 * pos * ts]
        else:
            for i in range(0, N1, 2):
                a, b = sm[i].copy(), sm[i + 1].copy()
                sm[i], sm[i + 1] = a + b, a - b
    A = np.zeros((N1, N2), complex)
    n2 = np.arange(N2)
    for pos in range(N1):
        k1 = pos_to_bin(pos, N1)
        A[k1] = sm[pos] * W[n2 * k1]
    return A


def rows(A, W, N1, N2, inv):
    """Two-level DFT of every row, as k_mix_rows does it: N2 = Na*Nb, n = Nb na + nb, k = ka + Na kb."""
    wm = W[np.arange(N2) * N1]
    if inv:
        wm = wm.conj()
    Na = max(a for a in range(1, int(N2 ** 0.5) + 1) if N2 % a == 0)
    Nb = N2 // Na
    out = np.zeros_like(A)
    for r in range(A.shape[0]):
        x = A[r]
        tb = np.zeros(N2, complex)
        for t in range(N2):
            ka, nb = divmod(t, Nb)
            acc = sum(x[Nb * na + nb] * wm[((na * ka) % Na) * Nb] for na in range(Na))
            tb[t] = acc * wm[nb * ka]
        for t in range(N2):
            ka, kb = divmod(t, Nb)
            out[r, ka + Na * kb] = sum(tb[ka * Nb + nb] * wm[((nb * kb) % Nb) * Na] for nb in range(Nb))
    return out


def cols_inv(Q, W, N1, lg, N2):
    sm = np.zeros((N1, N2), complex)
    n2 = np.arange(N2)
    for k1 in range(N1):
        sm[bin_to_pos(k1, N1)] = Q[k1] * W[n2 * k1].conj()
    for kind, q in reversed(stages(N1)):
        if kind == "r2":
            for i in range(0, N1, 2):
                a, b = sm[i].copy(), sm[i + 1].copy()
                sm[i], sm[i + 1] = a + b, a - b
        else:
            ts = N2 * (N1 // (4 * q))
            for g in range(0, N1, 4 * q):
                for pos in range(q):
                    i = g + pos
                    b0 = sm[i].copy()
                    b1 = sm[i + q] * W[pos * ts].conj()
                    b2 = sm[i + 2 * q] * W[2 * pos * ts].conj()
                    b3 = sm[i + 3 * q] * W[3 * pos * ts].conj()
                    t0, t1, t2, t3 = b0 + b2, b0 - b2, b1 + b3, (b1 - b3) * 1j
                    sm[i], sm[i + q], sm[i + 2 * q], sm[i + 3 * q] = t0 + t2, t1 + t3, t0 - t2, t1 - t3
    return sm.reshape(-1) / (N1 * N2)


def combine(Z, c1, c2, S, N1, N2):
    """Z [P][N1*N2] in [k1][k2] order -> Q [2][N] same order."""
    N = N1 * N2
    P = (S + 1) // 2
    Q = np.zeros((2, N), complex)
    for o in range(N):
        k1, k2 = divmod(o, N2)
        k = k1 + N1 * k2
        km = (N - k) % N
        if k > km:
            continue
        om = (km % N1) * N2 + km // N1
        m1 = m2 = tg = 0j
        for p in range(P):
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
        Q[0][o] = m1 + 1j * m2
        Q[1][o] = tg + 1j * it
        if om != o:
            Q[0][om] = m1.conjugate() + 1j * m2.conjugate()
            Q[1][om] = tg.conjugate() + 1j * it.conjugate()
    return Q


def model_mix(src, delays, fs):
    S, L = src.shape
    N1, lg, N2 = split_length(L)
    W = np.exp(-2j * np.pi * np.arange(L) / L)
    P = (S + 1) // 2
    Z = []
    for p in range(P):
        z = src[2 * p] + 1j * (src[2 * p + 1] if 2 * p + 1 < S else 0.0)
        A = cols_fwd(z.astype(complex), W, N1, lg, N2)
        Z.append(rows(A, W, N1, N2, False).reshape(-1) if N2 > 1 else A.reshape(-1))
    c1 = [delays[s][0] * fs / L for s in range(S)]
    c2 = [delays[s][1] * fs / L for s in range(S)]
    Q = combine(Z, c1, c2, S, N1, N2)
    outs = []
    for q in range(2):
        B = rows(Q[q].reshape(N1, N2), W, N1, N2, True) if N2 > 1 else Q[q].reshape(N1, N2)
        outs.append(cols_inv(B, W, N1, lg, N2))
    mix = np.stack([outs[0].real, outs[0].imag])
    tgt, itf = outs[1].real, outs[1].imag
    norm = np.max(np.abs(mix)) + 1e-9
    return mix / norm, tgt / norm, itf / norm


def main():
    rng = np.random.default_rng(0)
    fs = 16000.0
    for L, S in ((2000, 3), (1000, 4), (96, 1), (640, 2), (375, 5), (512, 3)):
        src = rng.standard_normal((S, L))
        angles = [90.0, 40.0, 130.0, 65.0, 155.0][:S]
        delays = [O.far_field_delays(a, 0.04, 343.0) for a in angles]
        # check the plain transform first
        N1, lg, N2 = split_length(L)
        W = np.exp(-2j * np.pi * np.arange(L) / L)
        z = src[0] + 1j * src[-1]
        X = rows(cols_fwd(z.astype(complex), W, N1, lg, N2), W, N1, N2, False)  # [k1][k2]
        ref = np.fft.fft(z)
        k = np.arange(N1)[:, None] + N1 * np.arange(N2)[None, :]
        e_fft = np.abs(X - ref[k]).max()
        mix, tgt, itf = model_mix(src, delays, fs)
        rmix, rtgt, ritf = O.mix_far_field(list(src), angles, 0.04, 343.0, fs)
        e = max(np.abs(mix - rmix).max(), np.abs(tgt - rtgt).max(), np.abs(itf - ritf).max())
        print(f"L={L} S={S} N1={N1} N2={N2}  fft err {e_fft:.2e}  mixer err {e:.2e}")
        assert e_fft < 1e-9 and e < 1e-10


if __name__ == "__main__":
    main()

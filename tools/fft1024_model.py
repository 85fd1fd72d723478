"""numpy models of the index algebra added in round 1b, checked against numpy.fft on the CPU (design aids, no GPU):

  * csrc/avz_opt1024.cu - a real 1024-point frame as ONE 512-point complex transform of (even, odd) sample pairs, in
    the lane layout of avz_fft512.cuh (lane (k1, h) holds lo bins k = k1 + 16 j + 128 h and, after the mirror shuffle,
    Z[512 - k]); the inverse packing; the hop-512 overlap-add with the head/tail exchange between warp runs.
  * csrc/avz_mixer.cu - the mixed-radix Stockham row transform (radices <= 8) and the chirp-z (Bluestein) transform
    for lengths without a 2^a x (<= 1024) split.

python tools/fft1024_model.py
"""
import numpy as np


# ------------------------------------------------------------------------------------------ 1024 = 2 x 512 (even/odd)
def lane_bins():
    """(lane, j) -> lo bin k of avz_fft512.cuh's spectrum layout."""
    return {(k1 + 16 * h, j): k1 + 16 * j + 128 * h for h in range(2) for k1 in range(16) for j in range(8)}


def analyse_1024(x, win):
    """One-sided spectrum (513 bins) of one frame the way k1024_* does it; equals rfft(x * win) / sum(win)."""
    z = (x * win)[0::2] + 1j * (x * win)[1::2]
    Z = np.fft.fft(z) / 1024.0                       # the kernels fold 1 / (2 sum w) = 1/1024 into the window table
    W = np.exp(-2j * np.pi * np.arange(1024) / 1024)
    Y = np.zeros(513, complex)
    for (_lane, _j), k in lane_bins().items():       # every lane: its 8 lo bins k and their mirrors 512 - k
        zk, m = Z[k], Z[(512 - k) % 512]
        A = zk + np.conj(m)
        T = W[k] * (-1j) * (zk - np.conj(m))
        Y[k] = A + T
        Y[512 - k] = np.conj(A - T)
    Y[256] = 2 * np.conj(Z[256])                     # lane 0: hi[0]
    return Y


def synthesise_1024(S):
    """irfft(S, 1024) * sum(w) from a one-sided spectrum via one inverse 512-point transform (k1024_apply)."""
    S = S.copy()
    S[0], S[512] = S[0].real, S[512].real            # c2r ignores Im(DC), Im(Nyquist)
    W = np.exp(-2j * np.pi * np.arange(1024) / 1024)
    Z = np.zeros(512, complex)
    for (_lane, _j), k in lane_bins().items():
        a, c = S[k], S[512 - k]
        E = a + np.conj(c)
        O = (a - np.conj(c)) * np.conj(W[k])
        Z[k] = E + 1j * O                            # hermitian_pack: v[j] = Sa + i Sb ...
        Z[(512 - k) % 512] = np.conj(E) + 1j * np.conj(O)   # ... and conj(Sa) + i conj(Sb) at the mirror lane
    s256 = S[256]
    Z[256] = 2 * s256.real + 1j * (-2 * s256.imag)   # ny = (Re Sa[256], Re Sb[256])
    z = np.fft.ifft(Z) * 512 * 0.5                   # unnormalised inverse, the 1/2 lives in the synthesis window
    out = np.empty(1024)
    out[0::2], out[1::2] = z.real, z.imag
    return out                                       # = irfft(S) * 512


def ola_runs(frames_win, n_warps=4):
    """Hop-512 overlap-add of windowed frames split into consecutive runs (one per warp): inside a run the open
    half-frame stays in registers, the first block of a later run is parked and completed from the previous run's tail."""
    T = len(frames_win)
    blocks = {}
    per = -(-T // n_warps)
    tails, parked = {}, {}
    for w in range(n_warps):
        fa, fb = w * per, min(T, (w + 1) * per)
        tail = None
        for t in range(fa, fb):
            head, new_tail = frames_win[t][:512], frames_win[t][512:]
            if t > 0:
                if t == fa:
                    parked[w] = (t, head)
                else:
                    blocks[t] = tail + head
            tail = new_tail
        if fa < fb:
            tails[w] = tail
    for w, (t, head) in parked.items():              # after the barrier
        blocks[t] = tails[w - 1] + head
    return np.concatenate([blocks[t] for t in range(1, T)])


# ------------------------------------------------------------------------------------------ mixer: row FFT, chirp-z
def row_plan(n2):
    plan, n = [], n2
    for r in (5, 7, 3, 8, 4, 2):
        while n % r == 0 and n > 1:
            plan.append(r)
            n //= r
    return plan if n == 1 and plan else None


def stockham(x, plan):
    """k_mix_rows_fft / row_stage<R>: natural order in and out."""
    n2 = len(x)
    wm = np.exp(-2j * np.pi * np.arange(n2) / n2)
    cur, ns = x.astype(complex), 1
    for r in plan:
        nb, tstep = n2 // r, n2 // (ns * r)
        out = np.zeros(n2, complex)
        for j in range(nb):
            k = j % ns
            v = [cur[j + t * nb] * (1 if (t == 0 or ns == 1) else wm[k * t * tstep]) for t in range(r)]
            for u in range(r):
                out[(j - k) * r + k + u * ns] = sum(v[t] * wm[((t * u) % r) * nb] for t in range(r))
        cur, ns = out, ns * r
    return cur


def chirp_z(u, m):
    """DFT of any length L as a cyclic convolution of length m >= 2L - 1 (avz_farfield_mix_f32, Bluestein path)."""
    L = len(u)
    n = np.arange(L, dtype=np.int64)
    c = np.exp(-1j * np.pi * ((n * n) % (2 * L)) / L)      # n^2 reduced mod 2L in integers
    kb = np.zeros(m, complex)
    kb[:L] = np.conj(c)
    kb[m - n[1:]] = np.conj(c[1:])
    a = np.zeros(m, complex)
    a[:L] = u * c
    conv = np.fft.ifft(np.fft.fft(a) * np.fft.fft(kb))
    return conv[:L] * c


def main():
    rng = np.random.default_rng(0)
    n = np.arange(1024)
    win = 0.5 - 0.5 * np.cos(2 * np.pi * n / 1024)
    x = rng.standard_normal(1024)
    Y = analyse_1024(x, win)
    ref = np.fft.rfft(x * win) / win.sum()
    assert np.abs(Y - ref).max() < 1e-13, "even/odd analysis"
    S = rng.standard_normal(513) + 1j * rng.standard_normal(513)
    assert np.abs(synthesise_1024(S) - np.fft.irfft(np.where(np.arange(513) % 512 == 0, S.real, S), 1024) * 512).max() < 1e-9
    # overlap-add: runs + exchange == plain loop
    T = 11
    fw = [rng.standard_normal(1024) for _ in range(T)]
    plain = np.zeros(512 * (T + 1))
    for t in range(T):
        plain[512 * t:512 * t + 1024] += fw[t]
    assert np.allclose(ola_runs(fw), plain[512:512 * T]), "head/tail exchange"
    for n2 in (125, 625, 375, 49, 16, 96, 1000, 7, 2):
        v = rng.standard_normal(n2) + 1j * rng.standard_normal(n2)
        assert np.abs(stockham(v, row_plan(n2)) - np.fft.fft(v)).max() < 1e-9 * n2, n2
    assert row_plan(823) is None and row_plan(11 * 5) is None
    for L, m in ((7, 16), (375, 1024), (4001, 8192), (12345, 25088)):
        v = rng.standard_normal(L) + 1j * rng.standard_normal(L)
        X = chirp_z(v, m)
        assert np.abs(X - np.fft.fft(v)).max() < 1e-9 * L, L
        assert np.abs(np.conj(chirp_z(np.conj(X), m)) / L - v).max() < 1e-9, L     # inverse = conj(DFT(conj Y)) / L
    print("fft1024_model: even/odd split, inverse packing, hop-512 overlap-add exchange, mixed-radix rows, chirp-z: ok")


if __name__ == "__main__":
    main()

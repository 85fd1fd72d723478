"""Lane-level numpy model of the register-resident 512-point warp FFT (csrc/avz_fft512.cuh).

32 lanes x 16 complex values.  Used to validate the index maps (transposition layout, half exchange,
mirror-lane unpack, inverse network) before they are written in CUDA.  Run: python tools/fft512_model.py
"""
import numpy as np

N = 512
W = lambda n, k: np.exp(-2j * np.pi * k / n)


def fft16(v, inv=False):
    return np.fft.ifft(v) * 16 if inv else np.fft.fft(v)


def shfl(regs, src):
    """regs: [32] values (one register across lanes); src: [32] source lane per lane."""
    return regs[src]


LANES = np.arange(32)
K1 = LANES & 15          # after the transposition a lane is (k1', h)
H = LANES >> 4
MIRROR = ((16 - K1) & 15) + 16 * (1 - H)
SIGN = np.where((LANES & 3) == 3, -1.0, 1.0)      # writer lanes n2 = 3 (mod 4) negate: rotates h=1 outputs by 8
TW = np.array([[W(512, L * k1) for k1 in range(16)] for L in range(32)]) * SIGN[:, None]


def forward(x):
    """x[512] -> lo[32][8], hi[32][8]: lane (k1',h): lo[j] = X[k1' + 16 j + 128 h], hi[j] = X[.. + 256]."""
    v = np.array([[x[32 * r + L] for r in range(16)] for L in range(32)])
    A = np.array([fft16(v[L]) for L in range(32)])
    Bm = A * TW                                        # [L=n2][k1]
    smem = {}
    for L in range(32):
        for k1 in range(16):
            smem[(k1, L)] = Bm[L, k1]
    u = np.array([[smem[(K1[l], 2 * m + H[l])] for m in range(16)] for l in range(32)])
    C = np.array([fft16(u[l]) for l in range(32)])     # h=1 lanes: C[q'] = C_1[(q'+8)%16] thanks to SIGN
    own = C[:, :8]
    recv = np.stack([shfl(C[:, 8 + j], LANES ^ 16) for j in range(8)], axis=1)
    E = np.where(H[:, None] == 1, recv, own)
    O = np.where(H[:, None] == 1, own, recv)
    t = O * np.array([[W(32, j + 8 * H[l]) for j in range(8)] for l in range(32)])
    return E + t, E - t


def bin_of(l, j, s):
    return K1[l] + 16 * j + 128 * H[l] + 256 * s


def mirror_values(lo, hi):
    """mir[l][j] = X[(512 - k) % 512] for k = bin_of(l, j, 0), via one shuffle per j (+ fix-up on k1'=0 lanes)."""
    recv = np.stack([shfl(hi[:, 7 - j], MIRROR) for j in range(8)], axis=1)
    mir = recv.copy()
    k0 = K1 == 0
    for j in range(1, 8):
        mir[k0, j] = recv[k0, j - 1]
    mir[k0, 0] = np.where(H[k0] == 1, hi[k0, 0], lo[k0, 0])
    return mir


def inverse(lo, hi):
    """Unnormalised inverse DFT from the (lo, hi) layout back to x[32 r + L] per lane L."""
    Ep = lo + hi
    Op = (lo - hi) * np.conj(np.array([[W(32, j + 8 * H[l]) for j in range(8)] for l in range(32)]))
    send = np.where(H[:, None] == 1, Ep, Op)
    keep = np.where(H[:, None] == 1, Op, Ep)
    recv = np.stack([shfl(send[:, j], LANES ^ 16) for j in range(8)], axis=1)
    C = np.concatenate([keep, recv], axis=1)           # h=0: C_0[q]; h=1: reg[q'] = C_1[(q'+8)%16]
    u = np.array([fft16(C[l], inv=True) for l in range(32)])
    smem = {}
    for l in range(32):
        for m in range(16):
            smem[(K1[l], 2 * m + H[l])] = u[l, m]
    Bm = np.array([[smem[(k1, L)] for k1 in range(16)] for L in range(32)]) * np.conj(TW)
    v = np.array([fft16(Bm[L], inv=True) for L in range(32)])
    x = np.zeros(512, complex)
    for L in range(32):
        for r in range(16):
            x[32 * r + L] = v[L, r]
    return x


def smem_pos(k1, n2):
    return k1 * 42 + (n2 & 1) * 24 + (n2 >> 1)


def check_banks():
    # forward write: lane L writes (k1, L) as 8-byte stores: half-warps must hit 16 distinct 8-byte banks
    for k1 in range(16):
        for half in (range(16), range(16, 32)):
            assert len({smem_pos(k1, L) % 16 for L in half}) == 16
    # forward read: lane (k1',h) reads 16 contiguous complex as 16-byte loads: quarter-warps, 8 distinct 16-byte banks
    for i in range(8):
        for q in range(4):
            lanes = range(8 * q, 8 * q + 8)
            units = set()
            for l in lanes:
                p = smem_pos(K1[l], 2 * (2 * i) + H[l])
                assert p % 2 == 0 and smem_pos(K1[l], 2 * (2 * i + 1) + H[l]) == p + 1
                units.add((p // 2) % 8)
            assert len(units) == 8
    print("bank checks ok; words per warp:", 16 * 42 * 2)


if __name__ == "__main__":
    rng = np.random.default_rng(0)
    x = rng.standard_normal(N) + 1j * rng.standard_normal(N)
    lo, hi = forward(x)
    X = np.fft.fft(x)
    err = max(abs(lo[l, j] - X[bin_of(l, j, 0)]) for l in range(32) for j in range(8))
    err = max(err, max(abs(hi[l, j] - X[bin_of(l, j, 1)]) for l in range(32) for j in range(8)))
    print("forward max err", err)
    mir = mirror_values(lo, hi)
    err = max(abs(mir[l, j] - X[(512 - bin_of(l, j, 0)) % 512]) for l in range(32) for j in range(8))
    print("mirror max err", err)
    xr = inverse(lo, hi) / N
    print("inverse max err", np.abs(xr - x).max())
    check_banks()


def hermitian_pack(Sa, Sb, Sa_ny, Sb_ny):
    """Two one-sided spectra held as lo-layout values (Sa[l][j] at bin_of(l,j,0) < 256, Nyquist separately)
    -> (lo, hi) of G = FFT(xa + i xb) for real xa, xb, so that one complex inverse yields both frames."""
    k0 = K1 == 0
    Sa = Sa.copy(); Sb = Sb.copy()
    dc = k0 & (H == 0)
    Sa[dc, 0] = Sa[dc, 0].real          # c2r ignores Im(DC)
    Sb[dc, 0] = Sb[dc, 0].real
    lo = Sa + 1j * Sb
    Gm = np.conj(Sa) + 1j * np.conj(Sb)                 # value of G at the mirrored bin 512 - k
    recv = np.stack([shfl(Gm[:, 7 - j], MIRROR) for j in range(8)], axis=1)
    hi = recv.copy()
    for j in range(1, 8):
        hi[k0, j] = recv[k0, j - 1]
    hi[k0 & (H == 1), 0] = Gm[k0 & (H == 1), 0]
    hi[dc, 0] = Sa_ny.real + 1j * Sb_ny.real            # Nyquist, Im ignored
    return lo, hi


def test_pack():
    rng = np.random.default_rng(1)
    Sa_full = rng.standard_normal(257) + 1j * rng.standard_normal(257)
    Sb_full = rng.standard_normal(257) + 1j * rng.standard_normal(257)
    Sa = np.array([[Sa_full[bin_of(l, j, 0)] for j in range(8)] for l in range(32)])
    Sb = np.array([[Sb_full[bin_of(l, j, 0)] for j in range(8)] for l in range(32)])
    lo, hi = hermitian_pack(Sa, Sb, Sa_full[256], Sb_full[256])
    z = inverse(lo, hi) / N
    xa = np.fft.irfft(Sa_full, 512)
    xb = np.fft.irfft(Sb_full, 512)
    print("pack+inverse err", np.abs(z.real - xa).max(), np.abs(z.imag - xb).max())


if __name__ == "__main__":
    test_pack()

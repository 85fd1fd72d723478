"""Memory-safety evidence without compute-sanitizer (closed on this pool): every output and workspace
buffer of the fused kernels is carved out of a larger allocation with poisoned guard bands on both sides; after the
kernels have run on ragged / odd / minimal shapes the bands must be untouched.  Plus argument checks of the C ABI."""
import ctypes as C

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

BAND = 4096          # bytes on each side
POISON = 0xA5


@pytest.fixture(scope="module")
def az():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    import avzoom
    avzoom._lib.load()
    return avzoom


class Guard:
    """Hands out tensors that live between two poisoned bands and checks the bands afterwards."""

    def __init__(self):
        self.bufs = []

    def like(self, t: torch.Tensor) -> torch.Tensor:
        n = t.numel() * t.element_size()
        pad = (-n) % 256
        raw = torch.full((BAND + n + pad + BAND,), POISON, dtype=torch.uint8, device=t.device)
        view = raw[BAND:BAND + n].view(t.dtype).view(t.shape)
        self.bufs.append((raw, n))
        return view

    def empty(self, shape, dtype) -> torch.Tensor:
        return self.like(torch.empty(shape, dtype=dtype, device="cuda"))

    def check(self):
        torch.cuda.synchronize()
        for i, (raw, n) in enumerate(self.bufs):
            lo, hi = raw[:BAND], raw[BAND + n:]
            assert bool((lo == POISON).all()), f"buffer {i}: bytes before the buffer were written"
            assert bool((hi == POISON).all()), f"buffer {i}: bytes after the buffer were written"


@pytest.mark.parametrize("preset,B,L", [("baseline_oracle", 3, 12345), ("baseline_oracle", 1, 512), ("baseline_oracle", 2, 513),
                                        ("oracle_debug", 3, 7001), ("oracle_debug", 2, 512), ("baseline_oracle", 37, 4000)])
@pytest.mark.parametrize("keep", [True, False])
def test_oracle_engine_stays_inside_its_buffers(az, preset, B, L, keep):
    from avzoom import pipeline
    cfg = az.PRESETS[preset]
    rng = np.random.default_rng(B * 100003 + L)
    tgt = rng.standard_normal((B, L)).astype(np.float32) * (rng.random((B, L)) < 0.6)      # gaps: a non-trivial IBM
    itf = rng.standard_normal((B, L)).astype(np.float32) * (rng.random((B, L)) < 0.6)
    mix = np.stack([tgt + itf, np.roll(tgt, 1, axis=1) + np.roll(itf, -1, axis=1)], axis=1).astype(np.float32)
    g = Guard()
    mix_d, tgt_d, itf_d = (g.like(torch.from_numpy(a).cuda()) for a in (mix, tgt, itf))
    mix_d.copy_(torch.from_numpy(mix)); tgt_d.copy_(torch.from_numpy(tgt)); itf_d.copy_(torch.from_numpy(itf))
    e = pipeline.OracleMvdr(cfg, B, L, mix_d.device, keep_spectrum=keep)
    ref = e.run(mix_d, tgt_d, itf_d).clone()
    for name in ("bits", "R", "msum", "w", "out", "peak", "ws", "spec"):
        t = getattr(e, name)
        if t is not None:
            setattr(e, name, g.like(t))
    e.peak.zero_()
    out = e.run(mix_d, tgt_d, itf_d)
    g.check()
    assert bool(torch.isfinite(ref).all()) and torch.equal(out, ref)     # and the guarded run computes the same bits


@pytest.mark.parametrize("n_fft,hop,B,L", [(1024, 512, 3, 32000), (1024, 512, 2, 1024), (1024, 512, 5, 1537), (512, 128, 3, 9999),
                                           (512, 256, 2, 777), (256, 64, 2, 1000)])
def test_learned_path_stays_inside_its_buffers(az, n_fft, hop, B, L):
    import dataclasses
    lib = az._lib.load()
    cfg = dataclasses.replace(az.PRESETS["full_audio"], n_fft=n_fft, hop=hop)
    rng = np.random.default_rng(0)
    F, T = cfg.n_freq, az.num_frames(L, n_fft, hop)
    g = Guard()
    mix = g.empty((B, 2, L), torch.float32)
    mix.copy_(torch.from_numpy(rng.standard_normal((B, 2, L)).astype(np.float32)))
    mask = g.empty((B, F, T), torch.float32)
    mask.copy_(torch.from_numpy(rng.uniform(0.05, 0.95, (B, F, T)).astype(np.float32)))
    p, st = az.ops._ptr, az.ops._stream
    X = g.empty((B, 2, F, T), torch.float32)
    az._lib.check(lib.avz_wave_features_f32(p(mix), B, L, n_fft, hop, 0, p(X), st()), "features")
    Xp = g.empty((B, F, T, 4), torch.float32)
    az._lib.check(lib.avz_wave_features_f32(p(mix), B, L, n_fft, hop, 2, p(Xp), st()), "features(physics)")
    Rp, ms = g.empty((B, F, 4), torch.float32), g.empty((B, F), torch.float32)
    ws = g.empty((max(int(lib.avz_ibm_cov_ws_bytes(B, L, n_fft, hop)), 4),), torch.uint8)
    nspec = int(lib.avz_spec_ws_bytes(B, L, n_fft, hop))
    w = g.empty((B, F, 2), torch.complex64)
    out, peak = g.empty((B, (T - 1) * hop), torch.float32), g.empty((B,), torch.float32)
    cc = cfg.to_c()
    d = az.steering_vectors(cfg, mix.device)
    outs = []
    for keep in ((True, False) if nspec > 0 else (False,)):
        spec = g.empty((nspec,), torch.uint8) if keep else None
        if keep:
            az._lib.check(lib.avz_wave_mask_cov_keep_f32(p(mix), p(mask), B, L, n_fft, hop, 0.0, 1e-6, p(Rp), p(ms), p(ws),
                                                         p(spec), st()), "cov keep")
        else:
            az._lib.check(lib.avz_wave_mask_cov_f32(p(mix), p(mask), B, L, n_fft, hop, 0.0, 1e-6, p(Rp), p(ms), p(ws), st()), "cov")
        az._lib.check(lib.avz_mvdr_weights_f32(p(Rp), p(d), B, F, C.byref(cc), p(w), st()), "weights")
        peak.zero_()
        if keep:
            az._lib.check(lib.avz_mvdr_apply_kept_f32(p(spec), p(w), p(None), p(mask), B, L, n_fft, hop, C.byref(cc), p(out),
                                                      p(peak), st()), "apply kept")
        else:
            az._lib.check(lib.avz_mvdr_apply_f32(p(mix), p(w), p(None), p(mask), B, L, n_fft, hop, C.byref(cc), p(out), p(peak),
                                                 st()), "apply")
        outs.append(out.clone())
    g.check()
    assert bool(torch.isfinite(outs[0]).all())
    if len(outs) == 2:
        assert torch.equal(outs[0], outs[1])


@pytest.mark.parametrize("R,Lrec", [(2, 40000), (3, 16001), (1, 70007)])
def test_chunk_kernels_stay_inside_their_buffers(az, R, Lrec):
    from avzoom.core import chunked
    lib = az._lib.load()
    cfg = az.PRESETS["full_audio"]
    win, stride = 32000, 16000
    n = chunked.n_windows(Lrec, win)
    B, F, T = R * n, 513, 64
    rng = np.random.default_rng(1)
    g = Guard()
    rec = g.empty((R, 2, Lrec), torch.float32)
    rec.copy_(torch.from_numpy(rng.standard_normal((R, 2, Lrec)).astype(np.float32)))
    mask = g.empty((B, F, T), torch.float32)
    mask.copy_(torch.from_numpy(rng.uniform(0.05, 0.95, (B, F, T)).astype(np.float32)))
    p, st = az.ops._ptr, az.ops._stream
    cv = az._lib.AvzChunkView(Lrec, n, stride)
    X = g.empty((B, 2, F, T), torch.float32)
    az._lib.check(lib.avz_chunk_features_f32(p(rec), R, C.byref(cv), win, 1024, 512, 0, p(X), st()), "chunk features")
    Rp, ms = g.empty((B, F, 4), torch.float32), g.empty((B, F), torch.float32)
    ws = g.empty((max(int(lib.avz_ibm_cov_ws_bytes(B, win, 1024, 512)), 4),), torch.uint8)
    spec = g.empty((int(lib.avz_spec_ws_bytes(B, win, 1024, 512)),), torch.uint8)
    az._lib.check(lib.avz_chunk_mask_cov_f32(p(rec), p(mask), R, C.byref(cv), win, 1024, 512, 0.0, 1e-6, p(Rp), p(ms), p(ws),
                                             p(spec), st()), "chunk cov")
    R64, ms64 = g.empty((B, F, 4), torch.float64), g.empty((B, F), torch.float64)
    ws64 = g.empty((int(lib.avz_wave_mask_cov_f64_ws_bytes(B, win, 1024, 512)),), torch.uint8)
    az._lib.check(lib.avz_chunk_mask_cov_f64(p(rec), p(mask), R, C.byref(cv), win, 1024, 512, 0.0, 1e-6, p(R64), p(ms64),
                                             p(ws64), st()), "chunk cov f64")
    assert float((Rp.double() - R64).abs().max() / R64.abs().max()) < 1e-5
    w = g.empty((B, F, 2), torch.complex64)
    cc = cfg.to_c()
    az._lib.check(lib.avz_mvdr_weights_f32(p(Rp), p(az.steering_vectors(cfg, rec.device)), B, F, C.byref(cc), p(w), st()), "w")
    outs = g.empty((B, 32256), torch.float32)
    res = []
    for sp in (spec, None):
        az._lib.check(lib.avz_chunk_mvdr_apply_f32(p(rec), p(sp), p(w), p(mask), R, C.byref(cv), win, 1024, 512, C.byref(cc),
                                                   p(outs), p(None), st()), "chunk apply")
        res.append(outs.clone())
    assert torch.equal(res[0], res[1])
    final, peak = g.empty((R, Lrec), torch.float32), g.empty((R,), torch.float32)
    peak.zero_()
    for use in (32000, 32256):
        az._lib.check(lib.avz_chunk_ola_f32(p(outs), R, n, 32256, Lrec, stride, use, p(final), p(peak), st()), "chunk ola")
    g.check()
    assert bool(torch.isfinite(final).all()) and torch.allclose(peak, final.abs().amax(dim=1))


def test_argument_checks(az):
    """Errors are codes + messages, never a launch with bad geometry (include/avzoom.h conventions)."""
    lib = az._lib.load()
    x = torch.zeros((1, 1, 300), dtype=torch.float32, device="cuda")
    Y = torch.zeros((1, 1, 257, 8), dtype=torch.complex64, device="cuda")
    p, st = az.ops._ptr, az.ops._stream
    # L < n_fft: scipy would shrink nperseg (the reference then indexes out of range) - rejected, with a message
    assert lib.avz_stft_f32(p(x), 1, 1, 300, 512, 128, p(Y), st()) == -1
    assert b"shorter than n_fft" in lib.avz_last_error()
    with pytest.raises(az._lib.AvzError, match="shorter than n_fft"):
        az.stft(np.zeros(300, np.float32), 512, 128)
    with pytest.raises(az._lib.AvzError):
        az.oracle_mask_mvdr(np.zeros((2, 300), np.float32), np.zeros(300, np.float32), np.zeros(300, np.float32))
    assert lib.avz_stft_f32(p(x), 1, 1, 300, 500, 125, p(Y), st()) == -1          # n_fft not supported
    assert lib.avz_stft_f32(p(x), 1, 1, 300, 256, 100, p(Y), st()) == -1          # hop does not divide n_fft
    assert lib.avz_stft_f32(p(None), 1, 1, 300, 256, 64, p(Y), st()) == -1        # null pointer
    assert lib.avz_ibm_cov_ws_bytes(70000, 4000, 512, 128) >= 0
    z = torch.zeros(16, dtype=torch.float32, device="cuda")
    assert lib.avz_ibm_cov_f32(p(z), p(z), p(z), 70000, 4000, 512, 128, 1e-6, p(z), p(z), p(z), p(z), st()) == -1   # B > 65535
    # shape mismatches are caught in the Python wrappers before any launch
    with pytest.raises(ValueError):
        az.masked_covariance(torch.zeros((2, 257, 8), dtype=torch.complex64, device="cuda"),
                             torch.zeros((257, 9), dtype=torch.float32, device="cuda"))
    with pytest.raises(ValueError):
        az.beamform(torch.zeros((100, 2), dtype=torch.complex64, device="cuda"),
                    torch.zeros((2, 257, 8), dtype=torch.complex64, device="cuda"))
    with pytest.raises(ValueError):
        az.mvdr_weights(torch.zeros((257, 4), dtype=torch.float32, device="cuda"),
                        torch.zeros((100, 2), dtype=torch.complex64, device="cuda"))


def test_unfused_ops_at_baseline_batch_size(az):
    """B * F exceeds 65535 (the limit of grid.y / grid.z) from B = 256 on: the unfused ops at B = 1024."""
    rng = np.random.default_rng(2)
    B, F, T = 1024, 257, 12
    Y = torch.from_numpy((rng.standard_normal((B, 2, F, T)) + 1j * rng.standard_normal((B, 2, F, T))).astype(np.complex64)).cuda()
    w = torch.from_numpy((rng.standard_normal((B, F, 2)) + 1j * rng.standard_normal((B, F, 2))).astype(np.complex64)).cuda()
    S = az.beamform(w, Y)
    want = (w[:, :, 0].conj()[:, :, None] * Y[:, 0] + w[:, :, 1].conj()[:, :, None] * Y[:, 1])
    assert float((S - want).abs().max()) < 1e-5
    m = az.geometric_mask(Y)
    assert m.shape == (B, F, T) and set(np.unique(m.cpu().numpy()).tolist()) <= {np.float32(0.01), np.float32(1.0)}
    X = az.logmag_ipd(Y)
    assert X.shape == (B, 2, F, T) and bool(torch.isfinite(X).all())
    assert torch.equal(X[1000], az.logmag_ipd(Y[1000]))
    nw = torch.from_numpy(rng.random((B, F, T)).astype(np.float32)).cuda()
    R = az.masked_covariance(Y, nw, packed=True)
    assert torch.equal(R[777], az.masked_covariance(Y[777], nw[777], packed=True))
    wts = az.mvdr_weights(R, az.steering_vectors(az.PRESETS["baseline_oracle"], Y.device), az.PRESETS["baseline_oracle"])
    assert wts.shape == (B, F, 2) and bool(torch.isfinite(torch.view_as_real(wts)).all())


@pytest.mark.parametrize("B,S,L", [(3, 4, 64000), (20, 3, 840), (2, 1, 96), (17, 4, 4096), (2, 2, 8)])
@pytest.mark.parametrize("misalign", [0, 1])
def test_far_field_mixer_stays_inside_its_buffers(az, B, S, L, misalign):
    """The cluster-resident mixer (float4 traffic when every signal is 16-byte aligned, 4-byte accesses otherwise:
    `misalign` shifts the sources and the outputs by one float) and the multi-pass mixer write nothing outside
    mix / tgt / itf / ws, and agree with each other."""
    from avzoom import _lib
    lib = _lib.load()
    rng = np.random.default_rng(B + 7 * S + L)
    src = torch.from_numpy(rng.standard_normal((B * S * L + 4,)).astype(np.float32)).cuda()[misalign:misalign + B * S * L]
    th = np.deg2rad([90.0, 40.0, 130.0, 65.0][:S])
    delays = np.ascontiguousarray(np.stack([0.02 * np.cos(th) / 343.0, 0.02 * np.cos(th - np.pi) / 343.0], axis=1))
    dp = delays.ctypes.data_as(C.POINTER(C.c_double))
    outs = {}
    for name, fn in (("cluster", lib.avz_farfield_mix_f32), ("passes", lib.avz_farfield_mix_passes_f32)):
        g = Guard()
        mix = g.empty((B * 2 * L + 4,), torch.float32)[misalign:misalign + B * 2 * L]
        tgt = g.empty((B * L + 4,), torch.float32)[misalign:misalign + B * L]
        itf = g.empty((B * L + 4,), torch.float32)[misalign:misalign + B * L]
        ws = g.empty((int(lib.avz_farfield_mix_ws_bytes(B, S, L)),), torch.uint8)
        _lib.check(fn(src.data_ptr(), dp, B, S, L, 16000.0, 1e-9, mix.data_ptr(), tgt.data_ptr(), itf.data_ptr(),
                      ws.data_ptr(), torch.cuda.current_stream().cuda_stream), name)
        g.check()
        # the slack floats around a shifted output are part of the allocation but not of the signal: untouched too
        for (raw, _), n in zip(g.bufs[:3], (B * 2 * L, B * L, B * L)):
            body = raw[BAND:BAND + 4 * (n + 4)].view(torch.float32)
            assert bool((body[:misalign].view(torch.uint8) == POISON).all())
            assert bool((body[misalign + n:].view(torch.uint8) == POISON).all())
        outs[name] = (mix.clone(), tgt.clone(), itf.clone())
        assert torch.isfinite(mix).all() and abs(float(mix.abs().max()) - 1.0) < 1e-5
    for u, v in zip(outs["cluster"], outs["passes"]):
        assert float((u - v).norm() / (v.norm() + 1e-30)) < 5e-6 or float(v.abs().max()) < 1e-6


def _shifted(t: torch.Tensor) -> torch.Tensor:
    """The same values in a contiguous view that starts one element (4 or 8 bytes) into a larger allocation."""
    big = torch.empty((t.numel() + 8,), dtype=t.dtype, device=t.device)
    v = big[1:1 + t.numel()].view(t.shape)
    v.copy_(t)
    return v


def test_ops_accept_buffers_that_are_only_element_aligned(az):
    """numpy hands the reference arrays of any alignment; a contiguous view that starts one element into a larger
    CUDA allocation must give the operators the same results as an aligned tensor (vector loads may only be used where
    the pointer allows them) - never a misaligned-address fault."""
    from avzoom import synth as S
    mix_h, tgt_h, itf_h = S.make_batch(11, 3, 0.5, 2)
    mix, tgt, itf = (torch.from_numpy(a).cuda() for a in (mix_h, tgt_h, itf_h))
    smix, stgt, sitf = _shifted(mix), _shifted(tgt), _shifted(itf)
    assert smix.data_ptr() % 16 != 0 and smix.is_contiguous()
    for preset in ("baseline_oracle", "oracle_debug"):
        cfg = az.PRESETS[preset]
        a = az.oracle_mask_mvdr(mix, tgt, itf, cfg)
        b = az.oracle_mask_mvdr(smix, stgt, sitf, cfg)
        assert torch.equal(a, b)
    T = az.num_frames(mix.shape[-1], 512, 128)
    mask = torch.rand((3, 257, T), device="cuda")
    a = az.learned_mask_mvdr(mix, mask, az.PRESETS["baseline_learned"])
    b = az.learned_mask_mvdr(smix, _shifted(mask), az.PRESETS["baseline_learned"])
    assert torch.equal(a, b)
    for n_fft, hop in ((512, 128), (1024, 512), (256, 64)):
        assert torch.equal(az.wave_features(mix, n_fft, hop), az.wave_features(smix, n_fft, hop))
        Y = az.stft(mix, n_fft, hop)
        assert torch.equal(Y, az.stft(smix, n_fft, hop))
        assert torch.equal(az.istft(Y[:, 0], n_fft, hop), az.istft(_shifted(Y[:, 0].contiguous()), n_fft, hop))
        assert torch.equal(az.logmag_ipd(Y), az.logmag_ipd(_shifted(Y)))
    T2 = az.num_frames(mix.shape[-1], 1024, 512)
    mask2 = torch.rand((3, 513, T2), device="cuda")
    cfg2 = az.PRESETS["full_audio"]
    assert torch.equal(az.learned_mask_mvdr(mix, mask2, cfg2), az.learned_mask_mvdr(smix, _shifted(mask2), cfg2))
    n = min(a.shape[-1], tgt.shape[-1])
    est = a[:, :n].contiguous()
    assert torch.equal(az.sir_scores(est, tgt[:, :n].contiguous(), itf[:, :n].contiguous()),
                       az.sir_scores(_shifted(est), _shifted(tgt[:, :n].contiguous()), _shifted(itf[:, :n].contiguous())))
    pcm = az.ops.float_to_pcm16(est)
    assert torch.equal(az.ops.pcm16_to_float(pcm), (pcm.float() / 32768.0))


def test_engine_and_chunk_driver_accept_element_aligned_buffers(az):
    """The same for the engine writing into a caller's shifted output buffer and for the chunk driver reading shifted
    recordings (the 1024 / 512 kernels load float2 pairs: a recording that is only 4-byte aligned must still work)."""
    from avzoom import pipeline, synth as S
    from avzoom.core import chunked
    mix_h, tgt_h, itf_h = S.make_batch(12, 2, 0.7, 3)
    mix, tgt, itf = (torch.from_numpy(a).cuda() for a in (mix_h, tgt_h, itf_h))
    cfg = az.PRESETS["baseline_oracle"]
    eng = pipeline.OracleMvdr(cfg, 2, mix.shape[-1], mix.device)
    ref = eng.run(mix, tgt, itf).clone()
    out = _shifted(torch.zeros_like(ref))
    got = eng.run(_shifted(mix), _shifted(tgt), _shifted(itf), out=out)
    assert torch.equal(got, ref)
    rec = torch.from_numpy(np.random.default_rng(3).standard_normal((3, 2, 40001)).astype(np.float32)).cuda() * 0.1
    model = lambda X: torch.sigmoid(X[:, 0] - X[:, 0].mean())          # noqa: E731  (pointwise stand-in mask model)
    enh = chunked.ChunkedEnhancer(az.PRESETS["full_audio"])
    a = enh(rec, model)
    b = enh(_shifted(rec), model)
    assert torch.equal(a, b)

"""GPU parity tests of the n_fft 1024 / hop 512 fast path (avz_opt1024.cu: the learned pipelines' STFT shape on the
register-resident 512-point transform), through the C ABI, against the float64 oracle."""
import dataclasses

import numpy as np
import pytest
import torch

import oracle as O
from test_gpu_parity import WAVE_TOL, rel_l2, synth, to_oracle_cfg

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def az():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    import avzoom
    avzoom._lib.load()
    return avzoom


# L: a 2 s chunk (T = 64), ragged (tail frame partly outside), odd (channel 1 not 8-byte aligned: scalar loads),
# and the shortest legal signal (T = 3)
@pytest.mark.parametrize("L", [32000, 20320, 20321, 1024])
def test_wave_features_1024(az, L):
    mix, _, _ = synth(3, 3, 2.0, 3)
    mix = np.ascontiguousarray(mix[:, :, :L])
    T = O.n_frames(L, 1024, 512)
    X = az.wave_features(torch.from_numpy(mix).cuda(), 1024, 512).cpu().numpy()
    Xw = az.wave_features(torch.from_numpy(mix).cuda(), 1024, 512, "logmag_ipd_wrapped").cpu().numpy()
    P = az.wave_features(torch.from_numpy(mix).cuda(), 1024, 512, "physics").cpu().numpy()
    assert X.shape == (3, 2, 513, T) and P.shape == (3, 513, T, 4)
    for b in range(3):
        Yref = O.stft_scipy(mix[b], 1024, 512)
        Xref = O.logmag_ipd(Yref)
        strong = np.abs(Yref).min(axis=0) > 1e-4 * np.abs(Yref).max()
        assert np.max(np.abs(X[b, 0] - Xref[0])[strong]) < 1e-3
        assert np.max(np.abs(X[b, 0] - Xref[0])) < 0.1
        d = (X[b, 1] - Xref[1]).astype(np.float64)
        assert np.max(np.abs((d + np.pi) % (2 * np.pi) - np.pi)[strong]) < 1e-3
        assert np.mean(np.abs(d[strong]) > 1.0) < 5e-3
        assert np.all(np.abs(Xw[b, 1]) <= np.pi + 1e-6)
        dw = (Xw[b, 1] - Xref[1]).astype(np.float64)
        assert np.max(np.abs((dw + np.pi) % (2 * np.pi) - np.pi)[strong]) < 1e-3
        Pref = O.physics_features(Yref, 1024)
        assert np.array_equal(P[b, ..., 3], Pref[..., 3])
        assert np.max(np.abs(P[b, ..., 0] - Pref[..., 0])[strong]) < 1e-3
        assert np.max(np.abs(P[b, ..., 1:3][strong] - Pref[..., 1:3][strong])) < 1e-3


@pytest.mark.parametrize("preset", ["full_audio", "tf_lite"])
@pytest.mark.parametrize("dur", [2.0, 1.27, 4.0, 0.064, 0.1])   # 0.064 s = one n_fft (T = 3), 0.1 s: T = 5
def test_learned_mask_mvdr_1024(az, preset, dur):
    """Learned-mask MVDR at 1024/512 (full_audio.../inference.py:88-117; tf_lite_version/inference.py:85-179):
    random target-probability masks against the float64 oracle, pieces and whole."""
    from avzoom import ops
    cfg = az.PRESETS[preset]
    mix, _, _ = synth(3, 3, dur, 3)
    if dur == 1.27:
        mix = np.ascontiguousarray(mix[:, :, :-1])          # odd length: the second channel starts 4-byte aligned
    B, L = mix.shape[0], mix.shape[-1]
    T = O.n_frames(L, 1024, 512)
    rng = np.random.default_rng(int(dur * 100))
    mask = rng.random((B, 513, T)).astype(np.float32)
    mask[0, 9, :] = 1.0                                     # empty noise weight in one bin
    mask[1, :, min(5, T - 1)] = 0.0
    mix_d, mask_d = torch.from_numpy(mix).cuda(), torch.from_numpy(mask).cuda()
    # covariance
    Rp, _ = ops.wave_masked_covariance(mix_d, mask_d, cfg, None)
    Rp = Rp.cpu().numpy()
    for b in range(B):
        Yref = O.stft_scipy(mix[b].astype(np.float64), 1024, 512)
        Rref = O.masked_covariance_vec(Yref, 1.0 - mask[b].astype(np.float64), cfg.sqrt_eps, cfg.norm_eps)
        got = Rp[b]
        packed = np.stack([Rref[:, 0, 0].real, Rref[:, 1, 1].real, Rref[:, 0, 1].real, Rref[:, 0, 1].imag], axis=-1)
        assert got.shape == packed.shape and rel_l2(got, packed) < 2e-5
    # whole path
    out = az.learned_mask_mvdr(mix, mask, cfg)
    assert out.shape == (B, (T - 1) * 512)
    ocfg = to_oracle_cfg(cfg)
    for b in range(B):
        ref = O.learned_mask_mvdr_chunk(mix[b].T.astype(np.float64), lambda X: mask[b], ocfg)
        ref = O.peak_normalise(ref, ocfg.peak_eps)
        assert rel_l2(out[b], ref) < WAVE_TOL
    # reruns are bit-identical (fixed-order reductions, no float atomics)
    assert np.array_equal(az.learned_mask_mvdr(mix, mask, cfg), out)
    # kept spectrum (pass A writes both one-sided spectra, pass B starts from them) == recompute, bit for bit
    res = []
    for keep in (True, False):
        spec = ops.alloc_kept_spectrum(mix_d, cfg) if keep else None
        assert (spec is not None) == keep
        Rk, _ = ops.wave_masked_covariance(mix_d, mask_d, cfg, spec)
        w = ops.mvdr_weights(Rk, ops.steering_vectors(cfg, mix_d.device), cfg)
        o, pk = ops.mvdr_apply(mix_d, w, cfg, mask=mask_d if cfg.post in ("floor", "mask") else None, spec=spec)
        res.append((Rk.clone(), o.clone(), pk.clone()))
    assert torch.equal(res[0][0], res[1][0]) and torch.equal(res[0][1], res[1][1]) and torch.equal(res[0][2], res[1][2])


def test_apply_1024_post_modes(az):
    """Post-filter variants of the 1024/512 apply kernel: none / mask / floor (nb cell4: x M; inference.py:116)."""
    mix, _, _ = synth(3, 2, 2.0, 3)
    B, L = mix.shape[0], mix.shape[-1]
    T = O.n_frames(L, 1024, 512)
    mask = np.random.default_rng(3).random((B, 513, T)).astype(np.float32)
    for post in ("none", "mask", "floor"):
        cfg = dataclasses.replace(az.PRESETS["full_audio"], post=post, peak_eps=1e-9)
        out = az.learned_mask_mvdr(mix, mask, cfg)
        ocfg = to_oracle_cfg(cfg)
        for b in range(B):
            ref = O.learned_mask_mvdr_chunk(mix[b].T.astype(np.float64), lambda X: mask[b], ocfg)
            ref = O.peak_normalise(ref, ocfg.peak_eps)
            assert rel_l2(out[b], ref) < WAVE_TOL

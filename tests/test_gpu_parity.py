"""GPU parity tests: the CUDA path (through the C ABI) against the float64 oracle on the same inputs.

Bars: IBM / mask outputs bit-exact; waveforms within 1e-4 relative L2 of the
numpy/scipy float64 path; SIR within 0.05 dB.  Run on the B200 box with `pytest -m gpu`.
"""
import os

import numpy as np
import pytest
import torch

import oracle as O

pytestmark = pytest.mark.gpu

WAVE_TOL = 1e-4   # relative L2, BASELINE.json north_star
SIR_TOL_DB = 0.05


@pytest.fixture(scope="module")
def az():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    import avzoom
    avzoom._lib.load()
    return avzoom


def rel_l2(a, b):
    a = np.asarray(a, dtype=np.complex128 if np.iscomplexobj(a) or np.iscomplexobj(b) else np.float64)
    b = np.asarray(b)
    return float(np.linalg.norm(a - b) / (np.linalg.norm(b) + 1e-300))


def to_oracle_cfg(cfg):
    return O.PathConfig(fs=cfg.fs, n_fft=cfg.n_fft, hop=cfg.hop, mic_dist=cfg.mic_dist, c=cfg.c, angle_deg=cfg.angle_deg,
                        sigma=cfg.sigma, hp_hz=cfg.hp_hz, hp_mode=cfg.hp_mode, sqrt_eps=cfg.sqrt_eps,
                        norm_eps=cfg.norm_eps, w_eps=cfg.w_eps, post=cfg.post, post_floor=cfg.post_floor,
                        peak_eps=cfg.peak_eps)


def synth(config_id, n_utt, dur, n_int, start=0):
    from avzoom import synth as S
    return S.make_batch(config_id, n_utt, dur, n_int, start=start)


# --------------------------------------------------------------------------------------------- STFT / iSTFT
@pytest.mark.parametrize("n_fft,hop,shape", [
    (512, 128, (2, 80000)), (512, 256, (2, 32000)), (1024, 512, (2, 32000)), (256, 64, (1, 5001)),
    (512, 128, (3, 2, 16000)), (512, 128, (3, 9999)), (512, 128, (700,)), (1024, 256, (2, 1024)),
])
def test_stft_matches_scipy(az, n_fft, hop, shape):
    rng = np.random.default_rng(sum(shape) + n_fft)
    x = rng.standard_normal(shape).astype(np.float32)
    Y = az.stft(x, n_fft, hop)
    ref = O.stft_scipy(x, n_fft, hop)
    assert Y.shape == ref.shape and Y.dtype == np.complex64
    assert rel_l2(Y, ref) < 1e-6


@pytest.mark.parametrize("n_fft,hop,T", [(512, 128, 626), (512, 256, 126), (1024, 512, 64), (256, 64, 17), (512, 128, 2)])
def test_istft_matches_scipy(az, n_fft, hop, T):
    rng = np.random.default_rng(T)
    F = n_fft // 2 + 1
    S = (rng.standard_normal((2, F, T)) + 1j * rng.standard_normal((2, F, T))).astype(np.complex64)
    x, peak = az.istft(S, n_fft, hop, return_peak=True)
    ref = O.istft_scipy(S, n_fft, hop)
    assert x.shape == ref.shape == (2, (T - 1) * hop)
    assert rel_l2(x, ref) < 1e-6
    assert np.allclose(peak, np.abs(x).max(axis=-1), rtol=0, atol=0)


def test_stft_istft_round_trip(az):
    rng = np.random.default_rng(5)
    x = torch.from_numpy(rng.standard_normal((4, 2, 64000)).astype(np.float32)).cuda()
    Y = az.stft(x, 512, 128)
    xr = az.istft(Y, 512, 128)
    assert xr.shape == (4, 2, 64000)
    assert float((xr - x).norm() / x.norm()) < 1e-6


def test_stft_edge_and_errors(az):
    with pytest.raises(az._lib.AvzError):
        az.stft(np.zeros((2, 100), np.float32), 512, 128)       # shorter than n_fft (scipy would shrink nperseg)
    with pytest.raises(az._lib.AvzError):
        az.stft(np.zeros((2, 4000), np.float32), 500, 125)      # unsupported n_fft
    Y = az.stft(np.zeros((2, 4000), np.float32), 512, 128)      # silence -> exact zeros
    assert not Y.any()


# --------------------------------------------------------------------------------------------- masks
def test_ibm_from_spectra_bit_exact(az):
    rng = np.random.default_rng(1)
    a = (rng.standard_normal((257, 300)) + 1j * rng.standard_normal((257, 300))).astype(np.complex64)
    b = (rng.standard_normal((257, 300)) + 1j * rng.standard_normal((257, 300))).astype(np.complex64)
    b[:10] = a[:10]                     # exact ties -> 0 in both polarities
    b[10:20] = 0
    a[15:20] = 0                        # 0 vs 0
    b[20, :50] = a[20, :50] * np.float32(1.0000001)
    got = az.ibm(a, b)                  # noise polarity: |S_int| > |S_tgt| with (S_tgt, S_int) = (a, b)
    ref = O.ibm_noise_mask(a.astype(np.complex128), b.astype(np.complex128))
    assert np.array_equal(got, ref.astype(np.float32))
    got_t = az.ibm_target_label(a, b)
    assert np.array_equal(got_t, O.ibm_target_label(a.astype(np.complex128), b.astype(np.complex128)))
    assert not np.any((got == 1) & (got_t == 1))


def test_geometric_mask_bit_exact(az, golden_dir):
    g = np.load(os.path.join(golden_dir, "ref_helpers.npz"))
    got = az.geometric_mask(g["geo_Y"])
    assert np.array_equal(got, g["geo_mask"].astype(np.float32))


# --------------------------------------------------------------------------------------------- pieces vs oracle
def test_masked_covariance_weights_beamform(az, golden_dir):
    g = np.load(os.path.join(golden_dir, "ref_helpers.npz"))
    Y, mask, f = g["bm_Y"], g["bm_mask"], g["asv_f_bins"]
    cfg = az.PRESETS["tf_lite"]
    R = az.masked_covariance(Y, 1.0 - mask, sqrt_eps=1e-10, norm_eps=1e-6)
    Rref = O.masked_covariance_vec(Y.astype(np.complex128), 1.0 - mask.astype(np.float64), 1e-10, 1e-6)
    assert R.shape == (513, 2, 2) and rel_l2(R, Rref) < 1e-6
    d = O.all_steering_vectors(f, 90.0, 0.04, 343.0)
    w = az.mvdr_weights(Rref.astype(np.complex64), d.astype(np.complex64), cfg)
    wref = O.mvdr_weights(Rref.astype(np.complex64).astype(np.complex128), d.astype(np.complex64), cfg.sigma, cfg.w_eps)
    assert rel_l2(w, wref) < 1e-6
    S = az.beamform(w, Y)
    assert rel_l2(S, O.beamform(w.astype(np.complex128), Y.astype(np.complex128))) < 1e-6
    # the reference's own batch_mvdr output (golden) through our three ops
    w2 = az.mvdr_weights(R, d.astype(np.complex64), cfg)
    assert rel_l2(az.beamform(w2, Y), g["bm_out"]) < 2e-5


def test_features(az):
    mix, _, _ = synth(3, 1, 2.0, 3)
    Yref = O.stft_scipy(mix[0], 1024, 512)
    Y = az.stft(mix[0], 1024, 512)
    X = az.logmag_ipd(Y)
    Xref = O.logmag_ipd(Yref)
    assert X.shape == Xref.shape == (2, 513, 64)
    strong = np.abs(Yref).min(axis=0) > 1e-4 * np.abs(Yref).max()
    assert np.max(np.abs(X[0] - Xref[0])[strong]) < 1e-3   # log / angle of weak bins amplify the f32 STFT noise
    assert np.max(np.abs(X[0] - Xref[0])) < 0.1
    # the un-wrapped IPD is discontinuous where an angle sits on the +-pi branch cut: there the float32 and
    # float64 spectra may pick opposite signs of pi, so values agree modulo 2 pi and jumps are rare
    d = (X[1] - Xref[1]).astype(np.float64)
    d_wrapped = np.abs((d + np.pi) % (2 * np.pi) - np.pi)
    assert np.max(d_wrapped[strong]) < 1e-3
    assert np.mean(np.abs(d[strong]) > 1.0) < 2e-3
    Xw = az.logmag_ipd(Y, wrapped=True)
    assert np.all(np.abs(Xw[1]) <= np.pi + 1e-6)
    dw = (Xw[1] - Xref[1]).astype(np.float64)
    assert np.max(np.abs((dw + np.pi) % (2 * np.pi) - np.pi)[strong]) < 1e-3
    P = az.physics_features(Y)
    Pref = O.physics_features(Yref, 1024)
    assert P.shape == Pref.shape == (513, 64, 4)
    assert np.array_equal(P[..., 3], Pref[..., 3])
    assert np.max(np.abs(P[..., 1:3][strong] - Pref[..., 1:3][strong])) < 1e-3


@pytest.mark.parametrize("n_fft,hop", [(512, 128), (1024, 512)])
def test_features_same_spectrum(az, n_fft, hop):
    """The feature arithmetic on its own: the SAME complex64 spectrum goes to avz_features_f32 and to the
    float64 oracle, so the float32 STFT noise of weak bins cancels and what is left is the error of |.|, log and atan2
    (`feature_values` in avz_common.cuh, shared by every feature kernel).  Bounds hold on every non-zero bin."""
    mix, _, _ = synth(3, 2, 2.0, 3)
    Y = az.stft(torch.from_numpy(mix).cuda(), n_fft, hop)                      # the GPU's own spectrum, complex64
    Yn = Y.cpu().numpy()
    # plus an adversarial block: 12 decades of magnitude, angles on the axes and hugging the +-pi branch cut
    rng = np.random.default_rng(11)
    F, T = Yn.shape[-2:]
    mag = 10.0 ** rng.uniform(-12, 0, (2, F, T))
    ang = rng.uniform(-np.pi, np.pi, (2, F, T))
    ang[:, :, 0] = np.pi - 1e-7
    ang[:, :, 1] = -np.pi + 1e-7
    ang[:, :, 2] = np.pi / 2
    adv = (mag * np.exp(1j * ang)).astype(np.complex64)
    adv[0, :, 3] = (-mag[0, :, 3]).astype(np.float32)                          # negative real axis, imaginary part +0
    adv[1, :, 4] = (1j * mag[1, :, 4]).astype(np.complex64)                    # imaginary axis
    Yall = np.concatenate([Yn, adv[None]], axis=0)
    X = az.logmag_ipd(torch.from_numpy(Yall).cuda()).cpu().numpy()
    ok = (np.abs(Yall) > 1e-30).all(axis=1)
    assert ok.mean() > 0.98
    for b in range(Yall.shape[0]):
        Xref = O.logmag_ipd(Yall[b].astype(np.complex128))
        assert np.max(np.abs(X[b, 0].astype(np.float64) - Xref[0])[ok[b]]) <= 2e-6
        assert np.max(np.abs(X[b, 1].astype(np.float64) - Xref[1])[ok[b]]) <= 1e-6
    P = az.physics_features(torch.from_numpy(Yall).cuda()).cpu().numpy()
    for b in range(Yall.shape[0]):
        Pref = O.physics_features(Yall[b].astype(np.complex128), n_fft)
        assert np.max(np.abs(P[b, ..., 0].astype(np.float64) - Pref[..., 0])[ok[b]]) <= 2e-6
        assert np.max(np.abs(P[b, ..., 1:3].astype(np.float64) - Pref[..., 1:3])[ok[b]]) <= 2e-6


def test_sir_scores_match_reference_golden(az, golden_dir):
    g = np.load(os.path.join(golden_dir, "ref_helpers.npz"))
    est, t, i = g["score_in"]
    sc = az.sir_scores(est, t, i)
    assert np.allclose(sc[:2], g["score_osinr_osir"], atol=1e-4, rtol=0)
    assert np.allclose(sc[2:], g["score_sdr_sir"], atol=1e-4, rtol=0)


# --------------------------------------------------------------------------------------------- fused oracle path
def _check_oracle_path(az, mix, tgt, itf, cfg, wave_tol=WAVE_TOL):
    ocfg = to_oracle_cfg(cfg)
    out, parts = az.oracle_mask_mvdr(torch.from_numpy(mix).cuda(), torch.from_numpy(tgt).cuda(),
                                     torch.from_numpy(itf).cuda(), cfg, return_parts=True)
    F = cfg.n_freq
    masks = az.unpack_ibm(parts["ibm_bits"], F).cpu().numpy()
    out = out.cpu().numpy()
    raw = parts["x_raw"].cpu().numpy()
    for b in range(mix.shape[0]):
        ref, rp = O.oracle_mask_mvdr(mix[b], tgt[b], itf[b], ocfg, return_parts=True)
        assert np.array_equal(masks[b], rp["mask_noise"].astype(np.float32)), f"IBM differs for utterance {b}"
        assert rel_l2(raw[b], rp["x_raw"]) < wave_tol
        assert rel_l2(out[b], ref) < wave_tol
        n = min(len(ref), tgt.shape[1])
        sir_ref = O.osinr_osir(ref[:n], tgt[b, :n], itf[b, :n])[1]
        sir_got = O.osinr_osir(out[b, :n], tgt[b, :n], itf[b, :n])[1]
        assert abs(sir_ref - sir_got) < SIR_TOL_DB


def test_oracle_path_config1(az):
    """BASELINE config 1: 5 s, 1 target + 2 interferers, n_fft 512 / hop 128."""
    mix, tgt, itf = synth(1, 2, 5.0, 2)
    _check_oracle_path(az, mix, tgt, itf, az.PRESETS["baseline_oracle"])


def test_oracle_path_config2_sample(az):
    """BASELINE config 2 shape (4 s, 3 interferers), a sample of utterances checked in full."""
    mix, tgt, itf = synth(2, 4, 4.0, 3)
    _check_oracle_path(az, mix, tgt, itf, az.PRESETS["baseline_oracle"])


def test_oracle_path_hop256_and_ragged(az):
    mix, tgt, itf = synth(7, 2, 1.37, 2)      # L = 21920: not a multiple of the hop
    _check_oracle_path(az, mix, tgt, itf, az.PRESETS["oracle_debug"])
    _check_oracle_path(az, mix, tgt, itf, az.PRESETS["baseline_oracle"])


def test_oracle_path_long_utterances(az):
    """Two 25 s utterances (T = 3126 frames, ~100 frame chunks per utterance: the chunking, warm-up and hand-off logic
    of both passes at a very different shape from the 4 s benchmark batch), plus a batch of one."""
    cfg = az.PRESETS["baseline_oracle"]
    mix, tgt, itf = synth(1, 2, 25.0, 2)
    _check_oracle_path(az, mix, tgt, itf, cfg)
    _check_oracle_path(az, mix[:1, :, :70001], tgt[:1, :70001], itf[:1, :70001], cfg)


def test_oracle_path_real_speech_golden(az, golden_dir):
    """The reference's oracle_debug.main() run unmodified on a real-speech excerpt (tests/golden)."""
    g = np.load(os.path.join(golden_dir, "ref_speech_excerpt.npz"))
    mix = (g["mix_pcm"].astype(np.float32) / 32768.0).T[None].copy()
    tgt = (g["tgt_pcm"].astype(np.float32) / 32768.0)[None].copy()
    itf = (g["int_pcm"].astype(np.float32) / 32768.0)[None].copy()
    cfg = az.PRESETS["oracle_debug"]
    _check_oracle_path(az, mix, tgt, itf, cfg)
    out = az.oracle_mask_mvdr(mix, tgt, itf, cfg)
    assert rel_l2(out[0], g["oracle_debug_out_f64read"]) < WAVE_TOL
    # silence-padded input: zero frames give an all-zero IBM row and do not poison the covariance
    mixz = np.concatenate([mix, np.zeros_like(mix)], axis=-1)
    tz = np.concatenate([tgt, np.zeros_like(tgt)], axis=-1)
    iz = np.concatenate([itf, np.zeros_like(itf)], axis=-1)
    _check_oracle_path(az, mixz, tz, iz, az.PRESETS["baseline_oracle"])


def test_learned_mask_path_golden(az, golden_dir):
    """process_chunk of the reference (full_audio.../inference.py:88-118) with the mask its seeded U-Net
    produced; reference run with float32 reads."""
    g = np.load(os.path.join(golden_dir, "ref_speech_excerpt.npz"))
    lc = np.load(os.path.join(golden_dir, "ref_learned_chunk.npz"))
    mix = (g["mix_pcm"].astype(np.float32) / 32768.0)[:32000].T[None].copy()
    cfg = az.PRESETS["full_audio"]
    out = az.learned_mask_mvdr(mix, lc["masks"][:1], cfg)
    assert out.shape == (1, 32256)
    ref = O.learned_mask_mvdr_chunk(mix[0].T, lambda X: lc["masks"][0], to_oracle_cfg(cfg))
    assert rel_l2(out[0], ref) < WAVE_TOL
    assert rel_l2(out[0], lc["chunk0_out"]) < WAVE_TOL


@pytest.mark.parametrize("preset,hop", [("baseline_learned", 128), ("baseline_learned", 256), ("tf_lite", 128)])
def test_learned_mask_fast_path_512(az, preset, hop):
    """Learned-mask MVDR on the n_fft 512 fast path (BASELINE config 3 shape): random target-probability masks,
    ragged length, against the float64 oracle; the kept-spectrum variant (mask re-laid (B,T,264) behind the spectrum)
    and the recomputing variant (mask read in the reference layout) give bit-identical waveforms."""
    import dataclasses
    from avzoom import ops
    cfg = dataclasses.replace(az.PRESETS[preset], n_fft=512, hop=hop)
    mix, _, _ = synth(3, 3, 1.27, 3)                       # L = 20320: not a multiple of the hop
    B, L = mix.shape[0], mix.shape[-1]
    T = O.n_frames(L, 512, hop)
    rng = np.random.default_rng(hop)
    mask = rng.random((B, 257, T)).astype(np.float32)
    mask[0, 7, :] = 1.0                                     # empty noise weight in one bin
    mask[1, :, 5] = 0.0
    out = az.learned_mask_mvdr(mix, mask, cfg)
    assert out.shape == (B, (T - 1) * hop)
    ocfg = to_oracle_cfg(cfg)
    for b in range(B):
        ref = O.learned_mask_mvdr_chunk(mix[b].T.astype(np.float64), lambda X: mask[b], ocfg)
        ref = O.peak_normalise(ref, ocfg.peak_eps)
        assert rel_l2(out[b], ref) < WAVE_TOL
    # kept spectrum (transposed mask) == recompute (reference mask layout), bit for bit
    mix_d, mask_d = torch.from_numpy(mix).cuda(), torch.from_numpy(mask).cuda()
    res = []
    for keep in (True, False):
        spec = ops.alloc_kept_spectrum(mix_d, cfg) if keep else None
        Rp, _ = ops.wave_masked_covariance(mix_d, mask_d, cfg, spec)
        w = ops.mvdr_weights(Rp, ops.steering_vectors(cfg, mix_d.device), cfg)
        o, pk = ops.mvdr_apply(mix_d, w, cfg, mask=mask_d if cfg.post in ("floor", "mask") else None, spec=spec)
        res.append((Rp.clone(), o.clone(), pk.clone()))
    assert res[0][0].shape == res[1][0].shape and torch.equal(res[0][0], res[1][0])
    assert torch.equal(res[0][1], res[1][1]) and torch.equal(res[0][2], res[1][2])
    # pass B on the transposed mask pass A staged in `spec` (no second transposition): same bits again
    if cfg.post in ("floor", "mask"):
        spec = ops.alloc_kept_spectrum(mix_d, cfg)
        Rp, _ = ops.wave_masked_covariance(mix_d, mask_d, cfg, spec)
        w = ops.mvdr_weights(Rp, ops.steering_vectors(cfg, mix_d.device), cfg)
        o, pk = ops.mvdr_apply(mix_d, w, cfg, mask=mask_d, spec=spec, mask_staged=True)
        assert torch.equal(o, res[0][1]) and torch.equal(pk, res[0][2])


def test_geometric_mask_mvdr_pieces(az, golden_dir):
    """masked_mvdr.main's arithmetic (masked_mvdr.py:76-128) assembled from the public ops.  sigma = 1e-7 on a
    near-rank-1 covariance amplifies STFT rounding ~1e4 times, so float64 inputs run on the float64 operators like the
    reference's complex128 arrays - and then meet the 1e-4 budget with room to spare; the float32 operators on the same
    data are held to the distance the reference itself moves between float32 and float64 reads."""
    g = np.load(os.path.join(golden_dir, "ref_speech_excerpt.npz"))
    mix64 = (g["mix_pcm"].astype(np.float64) / 32768.0).T.copy()
    cfg = az.PRESETS["masked_mvdr"]
    ocfg = to_oracle_cfg(cfg)
    ref = O.geometric_mask_mvdr(mix64, ocfg)
    Yref = O.stft_scipy(mix64, cfg.n_fft, cfg.hop)
    Y = az.stft(torch.from_numpy(mix64).cuda(), cfg.n_fft, cfg.hop)
    assert Y.dtype == torch.complex128 and rel_l2(Y.cpu().numpy(), Yref) < 1e-13
    m = az.geometric_mask(Y)
    assert m.dtype == torch.float64
    assert np.mean(m.cpu().numpy() != O.geometric_phase_mask(Yref)) < 1e-5    # only exact angle ties can differ
    R = az.masked_covariance(Y, m, packed=True)
    w = az.mvdr_weights(R, az.steering_vectors(cfg, Y.device, wide=True), cfg)
    assert R.dtype == torch.float64 and w.dtype == torch.complex128
    x = az.istft(az.beamform(w, Y), cfg.n_fft, cfg.hop)
    x = az.peak_normalise(x[None], None, cfg.peak_eps)[0]
    assert rel_l2(x.cpu().numpy(), ref) < 1e-4
    assert rel_l2(x.cpu().numpy(), g["masked_mvdr_out_f64read"]) < 1e-4      # the reference's own output, float64 reads
    # float32 operators: same distance from the float64 result as the reference's own float32-read run
    ref_gap = rel_l2(g["masked_mvdr_out_f32read"], g["masked_mvdr_out_f64read"])
    Y32 = az.stft(torch.from_numpy(mix64.astype(np.float32)).cuda(), cfg.n_fft, cfg.hop)
    R32 = az.masked_covariance(Y32, az.geometric_mask(Y32), packed=True)
    w32 = az.mvdr_weights(R32, az.steering_vectors(cfg, Y32.device), cfg)
    x32 = az.istft(az.beamform(w32, Y32), cfg.n_fft, cfg.hop)
    x32 = x32 / (x32.abs().max() + 1e-6)
    assert rel_l2(x32.cpu().numpy(), ref) < max(5e-3, 3 * ref_gap)


def test_float64_operators_match_scipy(az):
    """avz_stft_f64 / avz_istft_f64 against scipy in float64 (the dtype scipy returns for float64 audio)."""
    rng = np.random.default_rng(5)
    for n_fft, hop, L in ((512, 128, 4001), (512, 256, 8000), (1024, 512, 32000), (256, 64, 1000)):
        x = rng.standard_normal((2, L))
        Yref = O.stft_scipy(x, n_fft, hop)
        Y = az.stft(x, n_fft, hop)
        assert Y.dtype == np.complex128 and Y.shape == Yref.shape
        assert rel_l2(Y, Yref) < 1e-13
        S = Yref[0] * (0.3 + 0.2j)
        assert rel_l2(az.istft(S, n_fft, hop), O.istft_scipy(S, n_fft, hop)) < 1e-13


@pytest.mark.parametrize("preset,B,dur", [("baseline_oracle", 5, 1.3), ("oracle_debug", 3, 2.0), ("baseline_oracle", 300, 0.25)])
def test_engine_kept_spectrum_equals_recompute(az, preset, B, dur):
    """The pre-allocated engine in its two modes - pass A keeps the packed mix spectrum and pass B streams it back
    through a TMA ring, or pass B recomputes the forward transform - must agree bit for bit, and with the ops path."""
    from avzoom import pipeline
    cfg = az.PRESETS[preset]
    mix, tgt, itf = synth(4, B, dur, 2)
    mix_d, tgt_d, itf_d = (torch.from_numpy(a).cuda() for a in (mix, tgt, itf))
    e_keep = pipeline.OracleMvdr(cfg, B, mix.shape[-1], mix_d.device, keep_spectrum=True)
    e_reco = pipeline.OracleMvdr(cfg, B, mix.shape[-1], mix_d.device, keep_spectrum=False)
    e_sep = pipeline.OracleMvdr(cfg, B, mix.shape[-1], mix_d.device, keep_spectrum=True, fused_norm=True)
    assert e_keep.spec is not None and e_reco.spec is None and not e_keep.fused_norm and e_sep.fused_norm
    def same(x, y):
        # an all-silent output is 0 / 0 = NaN after peak normalisation (as `s_out /= np.max(np.abs(s_out))` gives in
        # the reference); NaN patterns must agree, everything else bit for bit
        return torch.equal(torch.isnan(x), torch.isnan(y)) and torch.equal(torch.nan_to_num(x), torch.nan_to_num(y))

    a = e_keep.run(mix_d, tgt_d, itf_d).clone()
    b = e_reco.run(mix_d, tgt_d, itf_d).clone()
    assert same(a, b)
    assert same(a, e_sep.run(mix_d, tgt_d, itf_d))          # fused vs separate peak normalisation
    assert torch.equal(e_keep.bits, e_reco.bits) and torch.equal(e_keep.R, e_reco.R)
    assert torch.equal(e_keep.peak, e_reco.peak)
    c = az.oracle_mask_mvdr(mix_d, tgt_d, itf_d, cfg)
    assert same(a, c)
    assert same(e_keep.run(mix_d, tgt_d, itf_d), a)      # reruns are bit-stable
    ref = O.oracle_mask_mvdr(mix[0], tgt[0], itf[0], to_oracle_cfg(cfg))
    assert rel_l2(a[0].cpu().numpy(), ref) < WAVE_TOL


def test_streamed_engine_equals_single_engine(az):
    """StreamedOracleMvdr (consecutive batches on alternating CUDA streams / workspaces) returns, for every batch,
    exactly what one OracleMvdr returns; results stay valid until `depth` further submits."""
    from avzoom import pipeline
    cfg = az.PRESETS["baseline_oracle"]
    batches = []
    for i in range(5):
        mix, tgt, itf = synth(2, 6, 0.9, 3, start=10 * i)
        batches.append(tuple(torch.from_numpy(a).cuda() for a in (mix, tgt, itf)))
    L = batches[0][0].shape[-1]
    single = pipeline.OracleMvdr(cfg, 6, L, batches[0][0].device)
    want = [single.run(*b).clone() for b in batches]
    loop = pipeline.StreamedOracleMvdr(cfg, 6, L, batches[0][0].device, depth=2)
    got = []
    for k, b in enumerate(batches):
        out = loop.submit(*b)
        if k >= 1:                       # the previous result is still intact while this one is being computed
            loop.join()
            got.append(prev.clone())
        prev = out
    loop.join()
    got.append(prev.clone())
    assert len(got) == len(want)
    for a, b in zip(got, want):
        assert torch.equal(a, b)
    with pytest.raises(ValueError):
        pipeline.StreamedOracleMvdr(cfg, 6, L, batches[0][0].device, depth=0)


def test_ibm_bit_exact_at_scale(az):
    """IBM of the fused float32 path (near ties re-decided in float64) against the all-float64 GPU reference over
    ~16 M bins, the latter pinned to the CPU oracle on one utterance; plus degenerate inputs (identical references:
    every bin is an exact tie -> all zeros; the near-tie list overflows and the full recheck must still be right)."""
    cfg = az.PRESETS["baseline_oracle"]
    mix, tgt, itf = synth(2, 128, 4.0, 3, start=5000)
    mix_d, tgt_d, itf_d = (torch.from_numpy(a).cuda() for a in (mix, tgt, itf))
    bits, _, _ = az.ibm_covariance(mix_d, tgt_d, itf_d, cfg)
    ref_bits = az.ibm_exact_bits(tgt_d, itf_d, cfg)
    assert torch.equal(bits, ref_bits), f"{int((bits != ref_bits).sum())} IBM words differ"
    m = az.unpack_ibm(ref_bits[:1], cfg.n_freq).cpu().numpy()[0]
    mo = O.ibm_noise_mask(O.stft_scipy(tgt[0], 512, 128), O.stft_scipy(itf[0], 512, 128))
    assert np.array_equal(m, mo.astype(np.float32))
    # identical references
    bits2, _, _ = az.ibm_covariance(mix_d[:4], tgt_d[:4], tgt_d[:4].clone(), cfg)
    assert int(bits2.abs().sum()) == 0
    # scaled copy: |S_int| = 1.0000001 |S_tgt| everywhere they are non-zero -> all ones except exact-zero bins
    it2 = (tgt_d[:2].double() * 1.0000001).float()
    b3, _, _ = az.ibm_covariance(mix_d[:2], tgt_d[:2], it2, cfg)
    assert torch.equal(b3, az.ibm_exact_bits(tgt_d[:2], it2, cfg))


# --------------------------------------------------------------------------------------------- full-size properties
def test_config2_full_size_properties(az):
    """BASELINE config 2 at full size - 1024 DISTINCT 4 s utterances through the fused path: shape, per-utterance peak
    == 1, distortionless weights, permutation equivariance (utterances are independent - the sharding invariant: a
    shuffled batch gives the shuffled output bit for bit), and utterances spread over the batch against the float64
    oracle (IBM bit-exact, waveform, SIR)."""
    B = 1024
    from avzoom import synth as S
    mix_h, tgt_h, itf_h = S.make_batch(2, B, 4.0, 3, workers=max(1, min(16, os.cpu_count() or 1)))
    mix, tgt, itf = (torch.from_numpy(a).cuda() for a in (mix_h, tgt_h, itf_h))
    cfg = az.PRESETS["baseline_oracle"]
    out, parts = az.oracle_mask_mvdr(mix, tgt, itf, cfg, return_parts=True)
    assert out.shape == (B, 64000)
    assert torch.all(out.abs().amax(dim=1) == 1.0)
    perm = torch.from_numpy(np.random.default_rng(5).permutation(B)).cuda()
    out_p, parts_p = az.oracle_mask_mvdr(mix[perm].contiguous(), tgt[perm].contiguous(), itf[perm].contiguous(), cfg,
                                         return_parts=True)
    assert torch.equal(out_p, out[perm])
    assert torch.equal(parts_p["ibm_bits"], parts["ibm_bits"][perm])
    # d^H w = 1 above the high-pass (MVDR distortionless constraint), w = 0 below it
    d = az.steering_vectors(cfg, mix.device)
    resp = (d.conj()[None] * parts["w"]).sum(-1)
    hp = cfg.hp_bins()
    assert torch.all(parts["w"][:, :hp] == 0)
    assert float((resp[:, hp:] - 1).abs().max()) < 1e-5
    ocfg = to_oracle_cfg(cfg)
    picks = [0, 147, 148, 511, 777, 1023]          # first / last utterance, both sides of a wave of 148 SMs, the middle
    masks = az.unpack_ibm(parts["ibm_bits"][picks], cfg.n_freq).cpu().numpy()
    for i, b in enumerate(picks):
        ref, rp = O.oracle_mask_mvdr(mix_h[b], tgt_h[b], itf_h[b], ocfg, return_parts=True)
        assert np.array_equal(masks[i], rp["mask_noise"].astype(np.float32))
        got = out[b].cpu().numpy()
        assert rel_l2(got, ref) < WAVE_TOL
        n = min(len(ref), tgt_h.shape[1])
        assert abs(O.osinr_osir(ref[:n], tgt_h[b, :n], itf_h[b, :n])[1] - O.osinr_osir(got[:n], tgt_h[b, :n], itf_h[b, :n])[1]) < SIR_TOL_DB


@pytest.mark.parametrize("preset,B,dur", [("baseline_oracle", 5, 1.3), ("oracle_debug", 3, 2.0), ("baseline_oracle", 300, 0.25)])
def test_engine_kept_spectrum_equals_recompute(az, preset, B, dur):
    """The pre-allocated engine in its two modes - pass A keeps the packed mix spectrum and pass B streams it back
    through a TMA ring, or pass B recomputes the forward transform - must agree bit for bit, and with the ops path."""
    from avzoom import pipeline
    cfg = az.PRESETS[preset]
    mix, tgt, itf = synth(4, B, dur, 2)
    mix_d, tgt_d, itf_d = (torch.from_numpy(a).cuda() for a in (mix, tgt, itf))
    e_keep = pipeline.OracleMvdr(cfg, B, mix.shape[-1], mix_d.device, keep_spectrum=True)
    e_reco = pipeline.OracleMvdr(cfg, B, mix.shape[-1], mix_d.device, keep_spectrum=False)
    e_sep = pipeline.OracleMvdr(cfg, B, mix.shape[-1], mix_d.device, keep_spectrum=True, fused_norm=True)
    assert e_keep.spec is not None and e_reco.spec is None and not e_keep.fused_norm and e_sep.fused_norm
    def same(x, y):
        # an all-silent output is 0 / 0 = NaN after peak normalisation (as `s_out /= np.max(np.abs(s_out))` gives in
        # the reference); NaN patterns must agree, everything else bit for bit
        return torch.equal(torch.isnan(x), torch.isnan(y)) and torch.equal(torch.nan_to_num(x), torch.nan_to_num(y))

    a = e_keep.run(mix_d, tgt_d, itf_d).clone()
    b = e_reco.run(mix_d, tgt_d, itf_d).clone()
    assert same(a, b)
    assert same(a, e_sep.run(mix_d, tgt_d, itf_d))          # fused vs separate peak normalisation
    assert torch.equal(e_keep.bits, e_reco.bits) and torch.equal(e_keep.R, e_reco.R)
    assert torch.equal(e_keep.peak, e_reco.peak)
    c = az.oracle_mask_mvdr(mix_d, tgt_d, itf_d, cfg)
    assert same(a, c)
    assert same(e_keep.run(mix_d, tgt_d, itf_d), a)      # reruns are bit-stable
    ref = O.oracle_mask_mvdr(mix[0], tgt[0], itf[0], to_oracle_cfg(cfg))
    assert rel_l2(a[0].cpu().numpy(), ref) < WAVE_TOL


def test_streamed_engine_equals_single_engine(az):
    """StreamedOracleMvdr (consecutive batches on alternating CUDA streams / workspaces) returns, for every batch,
    exactly what one OracleMvdr returns; results stay valid until `depth` further submits."""
    from avzoom import pipeline
    cfg = az.PRESETS["baseline_oracle"]
    batches = []
    for i in range(5):
        mix, tgt, itf = synth(2, 6, 0.9, 3, start=10 * i)
        batches.append(tuple(torch.from_numpy(a).cuda() for a in (mix, tgt, itf)))
    L = batches[0][0].shape[-1]
    single = pipeline.OracleMvdr(cfg, 6, L, batches[0][0].device)
    want = [single.run(*b).clone() for b in batches]
    loop = pipeline.StreamedOracleMvdr(cfg, 6, L, batches[0][0].device, depth=2)
    got = []
    for k, b in enumerate(batches):
        out = loop.submit(*b)
        if k >= 1:                       # the previous result is still intact while this one is being computed
            loop.join()
            got.append(prev.clone())
        prev = out
    loop.join()
    got.append(prev.clone())
    assert len(got) == len(want)
    for a, b in zip(got, want):
        assert torch.equal(a, b)
    with pytest.raises(ValueError):
        pipeline.StreamedOracleMvdr(cfg, 6, L, batches[0][0].device, depth=0)


def test_ibm_bit_exact_at_scale(az):
    """IBM of the fused float32 path (near ties re-decided in float64) against the all-float64 GPU reference over
    ~16 M bins, the latter pinned to the CPU oracle on one utterance; plus degenerate inputs (identical references:
    every bin is an exact tie -> all zeros; the near-tie list overflows and the full recheck must still be right)."""
    cfg = az.PRESETS["baseline_oracle"]
    mix, tgt, itf = synth(2, 128, 4.0, 3, start=5000)
    mix_d, tgt_d, itf_d = (torch.from_numpy(a).cuda() for a in (mix, tgt, itf))
    bits, _, _ = az.ibm_covariance(mix_d, tgt_d, itf_d, cfg)
    ref_bits = az.ibm_exact_bits(tgt_d, itf_d, cfg)
    assert torch.equal(bits, ref_bits), f"{int((bits != ref_bits).sum())} IBM words differ"
    m = az.unpack_ibm(ref_bits[:1], cfg.n_freq).cpu().numpy()[0]
    mo = O.ibm_noise_mask(O.stft_scipy(tgt[0], 512, 128), O.stft_scipy(itf[0], 512, 128))
    assert np.array_equal(m, mo.astype(np.float32))
    # identical references
    bits2, _, _ = az.ibm_covariance(mix_d[:4], tgt_d[:4], tgt_d[:4].clone(), cfg)
    assert int(bits2.abs().sum()) == 0
    # scaled copy: |S_int| = 1.0000001 |S_tgt| everywhere they are non-zero -> all ones except exact-zero bins
    it2 = (tgt_d[:2].double() * 1.0000001).float()
    b3, _, _ = az.ibm_covariance(mix_d[:2], tgt_d[:2], it2, cfg)
    assert torch.equal(b3, az.ibm_exact_bits(tgt_d[:2], it2, cfg))


# --------------------------------------------------------------------------------------------- full-size properties
def test_config2_full_size_properties(az):
    """1024 x 4 s through the fused path: shape, per-utterance peak == 1, distortionless weights, and
    permutation equivariance (utterances are independent - the sharding invariant)."""
    B = 1024
    mix8, tgt8, itf8 = synth(2, 8, 4.0, 3)
    reps = B // 8
    mix = torch.from_numpy(np.tile(mix8, (reps, 1, 1))).cuda()
    tgt = torch.from_numpy(np.tile(tgt8, (reps, 1))).cuda()
    itf = torch.from_numpy(np.tile(itf8, (reps, 1))).cuda()
    cfg = az.PRESETS["baseline_oracle"]
    out, parts = az.oracle_mask_mvdr(mix, tgt, itf, cfg, return_parts=True)
    assert out.shape == (B, 64000)
    assert torch.all(out.abs().amax(dim=1) == 1.0)
    # copies of the same utterance anywhere in the batch give bit-identical results
    assert torch.equal(out[:8], out[-8:]) and torch.equal(out[8:16], out[:8])
    # d^H w = 1 above the high-pass (MVDR distortionless constraint), w = 0 below it
    d = az.steering_vectors(cfg, mix.device)
    resp = (d.conj()[None] * parts["w"]).sum(-1)
    hp = cfg.hp_bins()
    assert torch.all(parts["w"][:, :hp] == 0)
    assert float((resp[:, hp:] - 1).abs().max()) < 1e-5
    ref = O.oracle_mask_mvdr(mix8[3], tgt8[3], itf8[3], to_oracle_cfg(cfg))
    assert rel_l2(out[3].cpu().numpy(), ref) < WAVE_TOL


@pytest.mark.parametrize("preset,B,dur", [("baseline_oracle", 5, 1.3), ("oracle_debug", 3, 2.0), ("baseline_oracle", 64, 1.0),
                                          ("baseline_oracle", 300, 0.25), ("baseline_oracle", 1, 5.0)])
def test_fused_persistent_kernel_equals_separate_kernels(az, preset, B, dur):
    """avz_oracle_fused_f32 (pass A, weights, pass B and normalisation as tasks of one persistent kernel, spectrum ring
    in L2) against the five separate launches: same IBM bits; the same waveform bit for bit when both cut an utterance
    into the same 32-frame chunks (small batches), else to the float32 summation order of the covariance; reruns
    bit-identical (no float atomics, fixed reduction orders, whatever order the tasks were dequeued in)."""
    from avzoom import pipeline
    cfg = az.PRESETS[preset]
    mix, tgt, itf = synth(4, min(B, 16), dur, 2)
    rep = (B + mix.shape[0] - 1) // mix.shape[0]
    mix, tgt, itf = (np.tile(a, (rep,) + (1,) * (a.ndim - 1))[:B].copy() for a in (mix, tgt, itf))
    scale = np.linspace(0.5, 1.5, B, dtype=np.float32)
    mix, tgt, itf = mix * scale[:, None, None], tgt * scale[:, None], itf * scale[:, None]
    mix_d, tgt_d, itf_d = (torch.from_numpy(a).cuda() for a in (mix, tgt, itf))
    L = mix.shape[-1]
    sep = pipeline.OracleMvdr(cfg, B, L, mix_d.device)
    fus = pipeline.OracleMvdr(cfg, B, L, mix_d.device, fused=True)
    a = sep.run(mix_d, tgt_d, itf_d).clone()
    b = fus.run(mix_d, tgt_d, itf_d).clone()
    assert torch.equal(sep.bits, fus.bits)
    assert bool(torch.isfinite(b).all())
    assert rel_l2(fus.R.cpu().numpy(), sep.R.cpu().numpy()) < 1e-6
    assert rel_l2(b.cpu().numpy(), a.cpu().numpy()) < 1e-6
    assert torch.allclose(fus.peak, sep.peak, rtol=1e-5)
    if B <= 16:        # both paths cut T frames into 32-frame chunks: identical partial sums, identical everything
        assert torch.equal(fus.R, sep.R) and torch.equal(fus.w, sep.w) and torch.equal(b, a)
    for _ in range(3):
        assert torch.equal(fus.run(mix_d, tgt_d, itf_d), b)
    # against the float64 oracle directly
    for i in (0, B - 1):
        ref = O.oracle_mask_mvdr(mix[i], tgt[i], itf[i], to_oracle_cfg(cfg))
        assert rel_l2(b[i].cpu().numpy(), ref) < 1e-4


def test_pass_a_with_folded_weights_is_bit_identical(az):
    """avz_ibm_cov_weights_keep_f32 (finalize + 2x2 solve done by the block that completes an utterance's last chunk)
    against avz_ibm_cov_keep_f32 + avz_mvdr_weights_f32: same R, msum, w and waveform, bit for bit."""
    from avzoom import pipeline
    for preset, B, dur in (("baseline_oracle", 1, 5.0), ("baseline_oracle", 40, 1.1), ("oracle_debug", 3, 2.0)):
        cfg = az.PRESETS[preset]
        mix, tgt, itf = synth(6, min(B, 8), dur, 2)
        rep = (B + mix.shape[0] - 1) // mix.shape[0]
        mix_d, tgt_d, itf_d = (torch.from_numpy(np.tile(a, (rep,) + (1,) * (a.ndim - 1))[:B].copy()).cuda() for a in (mix, tgt, itf))
        a = pipeline.OracleMvdr(cfg, B, mix.shape[-1], mix_d.device)
        b = pipeline.OracleMvdr(cfg, B, mix.shape[-1], mix_d.device, fold_weights=True)
        assert b.fold_weights and not a.fold_weights
        xa, xb = a.run(mix_d, tgt_d, itf_d), b.run(mix_d, tgt_d, itf_d)
        assert torch.equal(a.R, b.R) and torch.equal(a.msum, b.msum) and torch.equal(a.w, b.w) and torch.equal(xa, xb)


def test_sparse_kept_spectrum_is_bit_identical_and_mismatch_is_loud(az):
    """Oracle post-filter: pass A keeps only the bins whose noise bit is clear (compacted), pass B reads them back:
    same waveform bit for bit as the dense kept spectrum and as the recomputing pass B; a dense spectrum read as sparse
    (or the reverse) gives NaN, not garbage."""
    from avzoom import pipeline
    for preset, B, dur in (("baseline_oracle", 6, 1.3), ("oracle_debug", 3, 2.0), ("baseline_oracle", 70, 0.6), ("baseline_oracle", 1, 5.0)):
        cfg = az.PRESETS[preset]
        mix, tgt, itf = synth(8, min(B, 8), dur, 3)
        rep = (B + mix.shape[0] - 1) // mix.shape[0]
        mix, tgt, itf = (np.tile(a, (rep,) + (1,) * (a.ndim - 1))[:B].copy() for a in (mix, tgt, itf))
        tgt[0, :4000] = 0.0                               # a stretch where every bin is noise-dominated: empty frames
        itf[-1, -4000:] = 0.0                             # ... and one where none is: full frames
        mix_d, tgt_d, itf_d = (torch.from_numpy(a).cuda() for a in (mix, tgt, itf))
        L = mix.shape[-1]
        dense = pipeline.OracleMvdr(cfg, B, L, mix_d.device)
        sparse = pipeline.OracleMvdr(cfg, B, L, mix_d.device, sparse_spectrum=True)
        reco = pipeline.OracleMvdr(cfg, B, L, mix_d.device, keep_spectrum=False)
        assert sparse.sparse and not dense.sparse
        a = dense.run(mix_d, tgt_d, itf_d).clone()
        b = sparse.run(mix_d, tgt_d, itf_d).clone()
        c = reco.run(mix_d, tgt_d, itf_d).clone()
        assert bool(torch.isfinite(a).all()) and torch.equal(a, b) and torch.equal(a, c)
        assert torch.equal(dense.R, sparse.R) and torch.equal(dense.bits, sparse.bits) and torch.equal(dense.peak, sparse.peak)
        # the ops-level pair
        spec = az.alloc_kept_spectrum(mix_d, cfg, ibm=True)
        bits, Rp, _ = az.ibm_covariance(mix_d, tgt_d, itf_d, cfg, spec, sparse=True)
        wts = az.mvdr_weights(Rp, az.steering_vectors(cfg, mix_d.device), cfg)
        x, pk = az.mvdr_apply(mix_d, wts, cfg, ibm_bits=bits, spec=spec, sparse=True)
        if cfg.peak_eps is not None:
            az.peak_normalise(x, pk, cfg.peak_eps)
        assert torch.equal(x, a)
        # layout mismatch: sparse pass A, dense pass B
        sparse.pass_a(mix_d, tgt_d, itf_d)
        sparse.weights()
        sparse.sparse = False
        sparse.pass_b(mix_d)
        assert bool(torch.isnan(sparse.out).all())


def test_postmask_store_skipping_is_bit_identical_and_guarded(az):
    """avz_ibm_cov_keep_postmask_f32 leaves out the stores of kept-spectrum sectors whose two bins pass B will zero:
    same waveform bit for bit as the fully written spectrum and as the recomputing pass B, whatever finite values the
    buffer held before; a pass B without the 1 - noise-mask post-filter on such a buffer gives NaN."""
    import ctypes as C
    import dataclasses
    from avzoom import pipeline
    for preset, B, dur in (("baseline_oracle", 6, 1.3), ("oracle_debug", 3, 2.0), ("baseline_oracle", 70, 0.6)):
        cfg = az.PRESETS[preset]
        mix, tgt, itf = synth(9, min(B, 8), dur, 3)
        rep = (B + mix.shape[0] - 1) // mix.shape[0]
        mix_d, tgt_d, itf_d = (torch.from_numpy(np.tile(a, (rep,) + (1,) * (a.ndim - 1))[:B].copy()).cuda() for a in (mix, tgt, itf))
        L = mix.shape[-1]
        full = pipeline.OracleMvdr(cfg, B, L, mix_d.device, skip_masked_stores=False)
        skip = pipeline.OracleMvdr(cfg, B, L, mix_d.device)
        reco = pipeline.OracleMvdr(cfg, B, L, mix_d.device, keep_spectrum=False)
        assert skip.skip and not full.skip
        # whatever finite values the buffer held before must not matter
        skip.spec.view(torch.float32)[: (skip.spec.numel() // 4)].copy_(
            (torch.rand(skip.spec.numel() // 4, device="cuda") - 0.5) * 1e30)
        a, b, c = (e.run(mix_d, tgt_d, itf_d).clone() for e in (full, skip, reco))
        assert bool(torch.isfinite(a).all()) and torch.equal(a, b) and torch.equal(a, c)
        assert torch.equal(skip.run(mix_d, tgt_d, itf_d), a)        # and again on its own leftovers
        assert torch.equal(az.oracle_mask_mvdr(mix_d, tgt_d, itf_d, cfg), a)
        # the guard: the same buffer read by a pass B with no post-filter
        skip.pass_a(mix_d, tgt_d, itf_d)
        skip.weights()
        none_cfg = dataclasses.replace(cfg, post="none").to_c()
        out = torch.empty_like(skip.out)
        az._lib.check(skip.lib.avz_mvdr_apply_kept_f32(az.ops._ptr(skip.spec), az.ops._ptr(skip.w), az.ops._ptr(None),
                                                       az.ops._ptr(None), B, L, cfg.n_fft, cfg.hop, C.byref(none_cfg),
                                                       az.ops._ptr(out), az.ops._ptr(None), az.ops._stream()), "apply")
        assert bool(torch.isnan(out).all())

"""GPU tests of the reference-mirroring entry points (same names / signatures as the reference), against vectors the
reference's own code produced (tests/golden) and against the oracle."""
import os
import wave

import numpy as np
import pytest
import torch

import oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def az():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    import avzoom
    avzoom._lib.load()
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    return avzoom


def rel_l2(a, b):
    a = np.asarray(a, dtype=np.complex128 if np.iscomplexobj(a) or np.iscomplexobj(b) else np.float64)
    return float(np.linalg.norm(a - np.asarray(b)) / (np.linalg.norm(b) + 1e-300))


def _write_wav(path, pcm):
    pcm = np.asarray(pcm, dtype="<i2")
    with wave.open(path, "wb") as w:
        w.setnchannels(1 if pcm.ndim == 1 else pcm.shape[1])
        w.setsampwidth(2)
        w.setframerate(16000)
        w.writeframes(pcm.tobytes())


def test_batch_mvdr_dropin_matches_reference_output(az, golden_dir):
    from avzoom.core import batch_mvdr as bm
    g = np.load(os.path.join(golden_dir, "ref_helpers.npz"))
    f = g["asv_f_bins"]
    for ang, sig, key in ((90.0, 1e-5, "bm_out"), (40.0, 1e-3, "bm_out_40")):
        d = bm.get_all_steering_vectors(f, ang, 0.04, 343.0)
        S = bm.batch_mvdr(g["bm_Y"], g["bm_mask"], f, d, sig)
        assert isinstance(S, np.ndarray) and S.shape == (513, 64)
        assert rel_l2(S, g[key]) < 2e-5
        # complex128 in (what scipy's STFT of float64 audio is) -> float64 operators, complex128 out
        S64 = bm.batch_mvdr(g["bm_Y"].astype(np.complex128), g["bm_mask"].astype(np.float64), f, d, sig)
        assert S64.dtype == np.complex128 and rel_l2(S64, g[key]) < 1e-10


def test_score_dropins_match_reference_output(az, golden_dir):
    from avzoom.core import metrics
    g = np.load(os.path.join(golden_dir, "ref_helpers.npz"))
    est, t, i = g["score_in"]
    assert np.allclose(metrics.calculate_osnr_osir(est, t, i), g["score_osinr_osir"], atol=1e-4, rtol=0)
    assert np.allclose(metrics.calculate_metrics_manual(est, t, i), g["score_sdr_sir"], atol=1e-4, rtol=0)


def test_oracle_debug_and_masked_mvdr_main(az, golden_dir, tmp_path, monkeypatch):
    """The reference's `main()` entry points on WAV files, against what the reference itself wrote."""
    from avzoom.core import oracle_debug, masked_mvdr
    g = np.load(os.path.join(golden_dir, "ref_speech_excerpt.npz"))
    monkeypatch.chdir(tmp_path)
    d = tmp_path / oracle_debug.OUTDIR
    d.mkdir(parents=True)
    _write_wav(str(d / "mixture.wav"), g["mix_pcm"])
    _write_wav(str(d / "target_reference.wav"), g["tgt_pcm"])
    _write_wav(str(d / "interference_reference.wav"), g["int_pcm"])
    out = oracle_debug.main()
    assert rel_l2(out, g["oracle_debug_out_f64read"]) < 1e-4
    assert (d / "output_oracle.wav").exists()
    # error behaviour: missing files -> message and None, like the reference
    os.remove(d / "target_reference.wav")
    assert oracle_debug.main() is None
    world = tmp_path / "run" / "World_Outputs"
    world.mkdir(parents=True)
    assert masked_mvdr.main(str(world)) is None                      # no mixture_3_sources.wav yet
    assert masked_mvdr.main(str(tmp_path / "nope")) is None
    _write_wav(str(world / "mixture_3_sources.wav"), g["mix_pcm"])
    p = masked_mvdr.main(str(world))
    assert p.endswith("MVDR_Outputs/output_masked_mvdr.wav") and os.path.exists(p)
    from avzoom import wavio
    ref = g["masked_mvdr_out_f64read"]
    # the waveform main() computes (sigma = 1e-7: float64 operators), against the reference's own float64-read output
    x = masked_mvdr.enhance(g["mix_pcm"].astype(np.float64) / 32768.0)
    assert rel_l2(x, ref) < 1e-4
    # the file it wrote holds that waveform as PCM16: what is left is the quantisation step (uniform, 1/32767 wide)
    y, _ = wavio.read(p, dtype="float64")
    q = (1.0 / 32767.0) / np.sqrt(12.0) * np.sqrt(len(ref)) / np.linalg.norm(ref)
    assert rel_l2(y * (32768.0 / 32767.0), ref) < 1.5 * q + 1e-4


def test_main_deploy_chunk_path(az, golden_dir):
    """process_chunk / main_deploy with the masks the reference's seeded U-Net produced (tight), and with our
    port of that U-Net on the GPU end to end (loose: random-init net, float32 features)."""
    from avzoom.core import chunked, models
    g = np.load(os.path.join(golden_dir, "ref_speech_excerpt.npz"))
    lc = np.load(os.path.join(golden_dir, "ref_learned_chunk.npz"))
    L = int(lc["L"])
    mix = (g["mix_pcm"].astype(np.float32) / 32768.0)[:L]

    class Replay(torch.nn.Module):
        def __init__(self, masks):
            super().__init__()
            self.masks = torch.from_numpy(masks).cuda()

        def forward(self, X):
            return self.masks[:X.shape[0]]

    full = chunked.enhance_waveform(mix, Replay(lc["masks"]), az.PRESETS["full_audio"])
    assert full.shape == (L,)
    assert rel_l2(full, lc["main_deploy_out"]) < 1e-4
    out0 = chunked.process_chunk(mix[:32000], Replay(lc["masks"]))
    assert rel_l2(out0, lc["chunk0_out"]) < 1e-4
    torch.manual_seed(0)
    net = models.FreqPreservingUNet().eval().cuda()
    full2 = chunked.enhance_waveform(mix, net, az.PRESETS["full_audio"])
    assert rel_l2(full2, lc["main_deploy_out"]) < 2e-3
    o, t_inf, t_mvdr = chunked.process_chunk(mix[:32000], net, chunk_idx=0)   # resnet-variant signature
    assert o.shape == (32256,) and t_inf > 0 and t_mvdr > 0


def test_wave_features_equal_spectrum_features(az):
    from avzoom import synth
    mix, _, _ = synth.make_batch(3, 2, 2.0, 3)
    x = torch.from_numpy(mix).cuda()
    # n_fft 1024 / hop 256 runs the generic kernels on both sides: same transform, same bits
    a = az.wave_features(x, 1024, 256)
    b = az.logmag_ipd(az.stft(x, 1024, 256))
    assert a.shape == b.shape == (2, 2, 513, 126)
    assert torch.equal(a, b)
    p = az.wave_features(x, 1024, 256, "physics")
    assert torch.equal(p, az.physics_features(az.stft(x, 1024, 256)))
    # 1024 / 512 features come from the register-resident fast path (another transform): equal up to float32 noise
    a = az.wave_features(x, 1024, 512)
    Y = az.stft(x, 1024, 512)
    b = az.logmag_ipd(Y)
    assert a.shape == b.shape == (2, 2, 513, 64)
    strong = (Y.abs().amin(dim=1) > 1e-4 * Y.abs().amax()).cpu()
    assert float((a[:, 0] - b[:, 0]).abs().cpu()[strong].max()) < 1e-3
    d = (a[:, 1] - b[:, 1]).double().cpu()
    assert float((torch.remainder(d + np.pi, 2 * np.pi) - np.pi).abs()[strong].max()) < 1e-3


def test_fast_path_features_512(az):
    """n_fft 512 features come from the register-FFT kernel with shared-memory staged, coalesced stores; compare with
    the oracle (float64) on well-conditioned bins, and with the spectrum-based op."""
    from avzoom import synth
    mix, _, _ = synth.make_batch(3, 3, 1.1, 3)            # T = 139: ragged last tile of 32 frames
    x = torch.from_numpy(mix).cuda()
    X = az.wave_features(x, 512, 128).cpu().numpy()
    assert X.shape == (3, 2, 257, 139)
    for b in range(3):
        Yref = O.stft_scipy(mix[b], 512, 128)
        Xref = O.logmag_ipd(Yref)
        strong = np.abs(Yref).min(axis=0) > 1e-4 * np.abs(Yref).max()
        assert np.max(np.abs(X[b, 0] - Xref[0])[strong]) < 1e-3
        d = (X[b, 1] - Xref[1]).astype(np.float64)
        assert np.max(np.abs((d + np.pi) % (2 * np.pi) - np.pi)[strong]) < 1e-3
        assert np.mean(np.abs(d[strong]) > 1.0) < 2e-3
    Xw = az.wave_features(x, 512, 128, "logmag_ipd_wrapped")
    assert float(Xw[:, 1].abs().max()) <= np.pi + 1e-6
    Xh = az.wave_features(x, 512, 256)                    # hop 256 variant of the same kernel
    assert Xh.shape == (3, 2, 257, 70) and bool(torch.isfinite(Xh).all())


def test_oracle_reverb_main(az, golden_dir, tmp_path):
    """oracle_reverb.main(args): IBM covariance, MVDR with --sigma/--hp, soft (IRM) post-filter, against the oracle
    assembled from the same pieces."""
    import argparse
    from avzoom.core import oracle_reverb
    from avzoom import wavio
    g = np.load(os.path.join(golden_dir, "ref_speech_excerpt.npz"))
    _write_wav(str(tmp_path / "mixture_wpe.wav"), g["mix_pcm"])
    _write_wav(str(tmp_path / "target_reference.wav"), g["tgt_pcm"])
    _write_wav(str(tmp_path / "interference_reference.wav"), g["int_pcm"])
    p = oracle_reverb.main(argparse.Namespace(outdir=str(tmp_path), sigma=1e-3, hp=100.0))
    y, _ = wavio.read(p, dtype="float64")
    mix = (g["mix_pcm"].astype(np.float64) / 32768.0).T
    tgt = g["tgt_pcm"].astype(np.float64) / 32768.0
    itf = g["int_pcm"].astype(np.float64) / 32768.0
    cfg = O.PathConfig(hop=256, sigma=1e-3, post="none", peak_eps=None)
    Y, St, Si = (O.stft_scipy(s, 512, 256) for s in (mix, tgt, itf))
    S = O.mvdr_oracle._mvdr_from_noise_weight(Y, O.ibm_noise_mask(St, Si), cfg)
    soft = np.sqrt(np.abs(St) ** 2 / (np.abs(St) ** 2 + np.abs(Si) ** 2 + 1e-10))
    ref = O.istft_scipy(S * soft, 512, 256)
    ref = ref / (np.max(np.abs(ref)) + 1e-9)
    # the waveform main() computes, before it is quantised to PCM16
    x = oracle_reverb.enhance(mix.astype(np.float32), tgt.astype(np.float32), itf.astype(np.float32), 1e-3, 100.0)
    assert rel_l2(x, ref) < 1e-4
    q = (1.0 / 32767.0) / np.sqrt(12.0) * np.sqrt(len(ref)) / np.linalg.norm(ref)   # PCM16 step of the written file
    assert rel_l2(y * (32768.0 / 32767.0), ref) < 1.5 * q + 1e-4
    assert oracle_reverb.main(argparse.Namespace(outdir=str(tmp_path / "missing"), sigma=1e-3, hp=100.0)) is None


def test_hybrid_hard_null_dropin(az, golden_dir):
    """`hybrid_hard_null_bf` against the reference's own output (golden), and the fused chunk path (covariance from the
    waveform -> hybrid-null weights -> beamform * mask -> iSTFT) against the oracle assembled from the same blocks."""
    from avzoom.final_pipeline import inference as fpi
    from avzoom import synth
    g = np.load(os.path.join(golden_dir, "ref_helpers.npz"))
    # complex128 in (the reference's dtype) -> float64 operators
    S = fpi.hybrid_hard_null_bf(g["bm_Y"].astype(np.complex128), g["bm_mask"].astype(np.float64), g["asv_f_bins"])
    assert isinstance(S, np.ndarray) and S.shape == (513, 64) and S.dtype == np.complex128
    assert rel_l2(S, g["hn_out"]) < 1e-4
    assert rel_l2(S, g["hn_out"]) < 1e-9                              # in fact: float64 closed forms vs eigh / cond / solve
    bypass = int((g["asv_f_bins"] < 200).sum())
    assert np.array_equal(S[:bypass], g["bm_Y"][0, :bypass].astype(np.complex128))   # mic 0 passes below 200 Hz, exactly
    # complex64 in -> float32 I/O around the same float64 closed form: the eigenvector of a float32-rounded covariance
    S32 = fpi.hybrid_hard_null_bf(g["bm_Y"], g["bm_mask"], g["asv_f_bins"])
    assert S32.dtype == np.complex64 and rel_l2(S32, g["hn_out"]) < 1e-3
    # a bin above the bypass whose interference covariance is exactly zero: the reference raises LinAlgError
    d5 = np.load(os.path.join(golden_dir, "ref_chunk_drivers.npz"))
    Yn, mn = d5["hn_nan_Y"].astype(np.complex128), d5["hn_nan_mask"].astype(np.float64)
    assert str(d5["hn_nan_raised"]).startswith("LinAlgError")
    with pytest.raises(np.linalg.LinAlgError):
        fpi.hybrid_hard_null_bf(Yn, mn, g["asv_f_bins"])
    Sn = fpi.hybrid_hard_null_bf(Yn, mn, g["asv_f_bins"], degenerate="nan")
    assert np.isnan(Sn[40]).all() and not np.isnan(np.delete(Sn, 40, axis=0)).any()
    Sd = fpi.hybrid_hard_null_bf(Yn, mn, g["asv_f_bins"], degenerate="das")
    assert not np.isnan(Sd).any() and rel_l2(np.delete(Sd, 40, axis=0), np.delete(Sn, 40, axis=0)) == 0.0
    So = fpi.hybrid_hard_null_bf(Yn, d5["hn_ok_mask"].astype(np.float64), g["asv_f_bins"])
    assert rel_l2(So, d5["hn_ok_out"]) < 1e-9

    mix, _, _ = synth.make_batch(6, 2, 2.0, 2)
    rng = np.random.default_rng(3)
    mask = rng.uniform(0.05, 0.95, (2, 513, 64)).astype(np.float32)

    class Replay(torch.nn.Module):
        def forward(self, X):
            return torch.from_numpy(mask).cuda()[:X.shape[0]]

    out = fpi.enhance_chunks(torch.from_numpy(mix).cuda(), Replay()).cpu().numpy()
    f = np.fft.rfftfreq(1024, 1 / 16000.0)
    for b in range(2):
        Y = O.stft_scipy(mix[b], 1024, 512)
        ref = O.istft_scipy(O.hybrid_hard_null(Y, mask[b], f) * mask[b], 1024, 512)
        assert out[b].shape == ref.shape == (32256,)
        assert rel_l2(out[b], ref) < 1e-4


def test_final_pipeline_batch_run(az, tmp_path, monkeypatch):
    from avzoom.final_pipeline import batch_run, config
    from avzoom.core import models
    monkeypatch.setattr(config, "SIM_DIR", str(tmp_path / "sim"))
    monkeypatch.setattr(config, "RESULTS_DIR", str(tmp_path / "res"))
    torch.manual_seed(0)
    reports = batch_run.run_batch(2, start_idx=3, n_interferers=2, model=models.FreqPreservingUNet().eval().cuda())
    assert len(reports) == 2 and reports[0]["run"] == "batch_test_003"
    for r in reports:
        assert np.isfinite(r["SIR_out"]) and np.isfinite(r["SIR_in"])
    assert os.path.exists(tmp_path / "res" / "batch_test_004_results" / "batch_test_004_enhanced.wav")

"""Pins the CPU oracle (oracle/) to vectors produced by the reference's own code
(oracle/make_golden.py -> tests/golden/*.npz).  CPU only."""
import os

import numpy as np
import pytest

import oracle as O


def rel_l2(a, b):
    a = np.asarray(a)
    b = np.asarray(b)
    return np.linalg.norm(a - b) / (np.linalg.norm(b) + 1e-300)


@pytest.fixture(scope="module")
def helpers(golden_dir):
    return np.load(os.path.join(golden_dir, "ref_helpers.npz"))


@pytest.fixture(scope="module")
def speech(golden_dir):
    return np.load(os.path.join(golden_dir, "ref_speech_excerpt.npz"))


@pytest.fixture(scope="module")
def learned(golden_dir):
    return np.load(os.path.join(golden_dir, "ref_learned_chunk.npz"))


def test_reference_constants(helpers):
    # masked_mvdr.py:9-18
    assert helpers["constants_masked_mvdr"].tolist() == [16000, 0.01, 343.0, 90.0, 2, 1e-7, 512, 256]


def test_steering_vectors(helpers):
    for args, ref, ref2 in zip(helpers["sv_args"], helpers["sv_out"], helpers["sv_full_out"]):
        got = O.steering_vector(*args)
        assert np.array_equal(got, ref)
        assert np.array_equal(got, ref2)
    f = helpers["asv_f_bins"]
    assert np.array_equal(O.all_steering_vectors(f, 90.0, 0.04, 343.0)[:, :, None], helpers["asv_out"])
    assert np.array_equal(O.all_steering_vectors(f, 40.0, 0.04, 343.0)[:, :, None], helpers["asv_out_40"])


def test_geometric_mask_bit_exact(helpers):
    got = O.geometric_phase_mask(helpers["geo_Y"].astype(np.complex128))
    assert np.array_equal(got, helpers["geo_mask"])
    assert set(np.unique(got)) == {0.01, 1.0}


def test_batch_mvdr(helpers):
    f = helpers["asv_f_bins"]
    Y = helpers["bm_Y"].astype(np.complex128)
    m = helpers["bm_mask"].astype(np.float64)
    for ang, sig, key in ((90.0, 1e-5, "bm_out"), (40.0, 1e-3, "bm_out_40")):
        d = O.all_steering_vectors(f, ang, 0.04, 343.0)[:, :, None]
        got = O.batch_mvdr_vec(Y, m, f, d, sig)
        assert rel_l2(got, helpers[key]) < 1e-13
    # the loop form with the tf_lite constants is the same arithmetic
    cfg = O.PRESETS["tf_lite"]
    S = O.mvdr_oracle._mvdr_from_noise_weight(Y, 1.0 - m, cfg)
    assert rel_l2(S, helpers["bm_out"]) < 1e-12


def test_scores(helpers):
    est, t, i = helpers["score_in"].astype(np.float64)
    assert np.allclose(O.osinr_osir(est, t, i), helpers["score_osinr_osir"], rtol=0, atol=1e-12)
    assert np.allclose(O.sir_sdr_unit_output(est, t, i), helpers["score_sdr_sir"], rtol=0, atol=1e-12)
    # full_audio.../inference.py:77-85 returns the SIR twice; same projection, eps 1e-6 on the norms
    assert abs(O.sir_sdr_unit_output(est, t, i)[1] - helpers["score_full_manual"][0]) < 1e-6


def test_far_field_mixer_parts(helpers):
    got = np.array([O.far_field_delays(a, 0.04, 343.0) for a in helpers["ffd_angles"]])
    assert np.array_equal(got, helpers["ffd_out"])
    assert np.array_equal(O.fractional_delay(helpers["fd_in"], 3.3e-5, 16000), helpers["fd_out"])


def test_far_field_mixer_matches_reference_mix_and_save(golden_dir):
    """world_building.mix_and_save run unmodified on three PCM16 sources (oracle/make_golden.py section 4).  The
    reference reads float32 and numpy >= 2 then transforms in single precision, so its own output carries ~3e-8 of
    float32 FFT noise against the float64 restatement."""
    z = np.load(os.path.join(golden_dir, "ref_mixer.npz"))
    src = z["src_pcm"].astype(np.float64) / 32768.0
    d, c, fs = z["d_c_fs"]
    mix, tgt, itf = O.mix_far_field(list(src), list(z["angles"]), d, c, fs)
    assert np.abs(mix.T - z["mix"]).max() < 2e-7
    assert np.abs(tgt - z["tgt"]).max() < 2e-7
    assert np.abs(itf - z["itf"]).max() < 2e-7
    assert np.array_equal(O.fractional_delay(z["fd2_in"].astype(np.float64), -4.1e-5, 16000), z["fd2_out"])


@pytest.mark.parametrize("n_fft,hop,L", [(512, 128, 80000), (512, 256, 32000), (1024, 512, 32000), (512, 128, 1000),
                                           (512, 128, 64001), (256, 64, 777)])
def test_stft_restatement_matches_scipy(n_fft, hop, L):
    rng = np.random.default_rng(L)
    x = rng.standard_normal((2, L))
    Z = O.stft_scipy(x, n_fft, hop)
    assert Z.shape == (2, n_fft // 2 + 1, O.n_frames(L, n_fft, hop))
    assert rel_l2(O.stft_np(x, n_fft, hop), Z) < 1e-14
    S = Z[0] * (1 + 0.3j)                    # non-Hermitian-consistent DC/Nyquist imaginary parts
    xr = O.istft_scipy(S, n_fft, hop)
    assert xr.shape == ((Z.shape[-1] - 1) * hop,)
    assert rel_l2(O.istft_np(S, n_fft, hop), xr) < 1e-13
    # round trip (COLA): istft(stft(x)) == x on the first L samples
    assert rel_l2(O.istft_scipy(Z[1], n_fft, hop)[:L], x[1]) < 1e-13


def _pcm(a):
    return a.astype(np.float64) / 32768.0


def test_oracle_debug_main(speech):
    """oracle_debug.main() unmodified (hop 256) on the real-speech excerpt."""
    mix, tgt, itf = _pcm(speech["mix_pcm"]).T, _pcm(speech["tgt_pcm"]), _pcm(speech["int_pcm"])
    got = O.oracle_mask_mvdr(mix, tgt, itf, O.PRESETS["oracle_debug"])
    ref64 = speech["oracle_debug_out_f64read"]
    assert got.shape == ref64.shape
    assert rel_l2(got, ref64) < 1e-12
    # the reference as shipped reads float32 (complex64 STFT inside scipy): same thing to ~1e-6
    assert rel_l2(got, speech["oracle_debug_out_f32read"]) < 2e-5


def test_masked_mvdr_main(speech):
    mix = _pcm(speech["mix_pcm"]).T
    got = O.geometric_mask_mvdr(mix, O.PRESETS["masked_mvdr"])
    assert rel_l2(got, speech["masked_mvdr_out_f64read"]) < 1e-9
    assert rel_l2(got, speech["masked_mvdr_out_f32read"]) < 1e-2   # sigma=1e-7: ill-conditioned, f32 STFT moves it


def test_process_chunk_and_main_deploy(speech, learned):
    L = int(learned["L"])
    mix = _pcm(speech["mix_pcm"])[:L].astype(np.float32)    # the reference reads float32
    masks = learned["masks"]
    out0 = O.learned_mask_mvdr_chunk(mix[:32000], lambda X: masks[0], O.PRESETS["full_audio"])
    assert out0.shape == learned["chunk0_out"].shape == (32256,)
    assert rel_l2(out0, learned["chunk0_out"]) < 2e-5         # reference STFT is complex64 here
    it = iter(masks)
    full = O.chunked_enhance(mix, lambda X: next(it), O.PRESETS["full_audio"], win=32000)
    assert full.shape == learned["main_deploy_out"].shape == (L,)
    assert rel_l2(full, learned["main_deploy_out"]) < 2e-5


def test_hybrid_hard_null(helpers):
    # Final_pipeline/src/inference.py:28-98: the reference's own output on seeded inputs
    got = O.hybrid_hard_null(helpers["bm_Y"].astype(np.complex128), helpers["bm_mask"].astype(np.float64),
                             helpers["asv_f_bins"])
    assert got.shape == (513, 64)
    assert rel_l2(got, helpers["hn_out"]) < 1e-12


@pytest.fixture(scope="module")
def drivers(golden_dir):
    return np.load(os.path.join(golden_dir, "ref_chunk_drivers.npz"))


def test_tflite_era_chunk_drivers(speech, learned, drivers):
    """process_audio_file (tf_lite_version/inference.py:245-391) and enhance_audio (Final_pipeline/src/inference.py:
    144-238), run unmodified with their TFLite interpreter replaced by a replay of seeded masks."""
    L = int(learned["L"])
    mix = _pcm(speech["mix_pcm"])[:L]
    masks = learned["masks"]
    for bf, key, kw in (("batch_mvdr", "process_audio_file_out", dict(mic_dist=0.04)),
                        ("hybrid_null", "enhance_audio_out", dict(mic_dist=0.08))):
        it = iter(masks)
        got = O.chunked_enhance_clipped(mix, lambda lm, ipd: next(it), bf, **kw)
        assert got.shape == (L,)
        assert rel_l2(got, drivers[key + "_f64read"]) < 1e-9
        it = iter(masks)
        got32 = O.chunked_enhance_clipped(mix.astype(np.float32), lambda lm, ipd: next(it), bf, **kw)
        assert rel_l2(got32, drivers[key + "_f32read"]) < 2e-5   # as shipped the reference's STFT and einsum are complex64


def test_hybrid_hard_null_degenerate_bin(helpers, drivers):
    """An interference covariance that is exactly zero above the bypass: the reference raises (np.linalg.cond of a NaN
    matrix), and so does the restatement; without that bin both agree."""
    Y = drivers["hn_nan_Y"].astype(np.complex128)
    f = helpers["asv_f_bins"]
    assert str(drivers["hn_nan_raised"]).startswith("LinAlgError")
    with pytest.raises(np.linalg.LinAlgError), np.errstate(all="ignore"):
        O.hybrid_hard_null(Y, drivers["hn_nan_mask"].astype(np.float64), f)
    got = O.hybrid_hard_null(Y, drivers["hn_ok_mask"].astype(np.float64), f)
    assert rel_l2(got, drivers["hn_ok_out"]) < 1e-12

"""GPU tests of the on-device far-field mixer and the PCM16 wire-format converters (rank 4),
through the C ABI, against the float64 oracle and the reference's own output (tests/golden/ref_mixer.npz)."""
import os

import numpy as np
import pytest
import torch

import oracle as O

pytestmark = pytest.mark.gpu

ANGLES = [90.0, 40.0, 130.0, 65.0, 155.0, 20.0, 110.0, 75.0]


@pytest.fixture(scope="module")
def az():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    import avzoom
    avzoom._lib.load()
    return avzoom


def rel_l2(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a - b) / (np.linalg.norm(b) + 1e-300))


# float32 transform of length L against float64: tolerance on rel-L2 of every output signal
MIX_TOL = 5e-6


@pytest.mark.parametrize("B,S,L", [(3, 4, 64000), (1, 3, 80000), (2, 1, 16000), (2, 2, 4096), (1, 5, 1000),
                                     (2, 8, 32000), (1, 3, 375), (1, 4, 96), (1, 2, 512)])
def test_mixer_matches_oracle(az, B, S, L):
    rng = np.random.default_rng(1000 * S + L % 997)
    src = rng.standard_normal((B, S, L)).astype(np.float32)
    delays = [O.far_field_delays(a, 0.04, 343.0) for a in ANGLES[:S]]
    mix, tgt, itf = az.ops.far_field_mix(torch.from_numpy(src).cuda(), delays, 16000.0)
    assert mix.shape == (B, 2, L) and tgt.shape == (B, L) and itf.shape == (B, L)
    for b in range(B):
        rm, rt, ri = O.mix_far_field(list(src[b].astype(np.float64)), ANGLES[:S], 0.04, 343.0, 16000.0)
        assert rel_l2(mix[b].cpu().numpy(), rm) < MIX_TOL
        assert rel_l2(tgt[b].cpu().numpy(), rt) < MIX_TOL
        if S > 1:
            assert rel_l2(itf[b].cpu().numpy(), ri) < MIX_TOL
        else:
            assert np.abs(itf[b].cpu().numpy()).max() < 1e-6
        assert abs(float(mix[b].abs().max()) - 1.0) < 1e-6


def test_mixer_matches_reference_mix_and_save(az, golden_dir):
    z = np.load(os.path.join(golden_dir, "ref_mixer.npz"))
    from avzoom.core import world_building as wb
    d, c, fs = z["d_c_fs"]
    assert (wb.D, wb.C, wb.FS) == (d, c, fs)
    assert [wb.ANGLE_TARGET, wb.ANGLE_INTERFERER_A, wb.ANGLE_INTERFERER_B] == list(z["angles"])
    src = az.ops.pcm16_to_float(torch.from_numpy(z["src_pcm"].astype(np.int16)).cuda())   # sf.read float32
    assert np.array_equal(src.cpu().numpy(), z["src_pcm"].astype(np.float32) / 32768.0)
    mix, tgt, itf = wb.mix_sources(src)
    assert rel_l2(mix.cpu().numpy().T, z["mix"]) < MIX_TOL
    assert rel_l2(tgt.cpu().numpy(), z["tgt"]) < MIX_TOL
    assert rel_l2(itf.cpu().numpy(), z["itf"]) < MIX_TOL
    # numpy in -> numpy out, single utterance
    m2, _, _ = wb.mix_sources(src.cpu().numpy())
    assert isinstance(m2, np.ndarray) and np.array_equal(m2, mix.cpu().numpy())


def test_apply_frac_delay_dropin(az, golden_dir):
    z = np.load(os.path.join(golden_dir, "ref_mixer.npz"))
    from avzoom.core import world_building as wb
    got = wb.apply_frac_delay(z["fd2_in"], -4.1e-5, 16000)
    assert isinstance(got, np.ndarray) and got.shape == z["fd2_out"].shape
    assert rel_l2(got, z["fd2_out"]) < MIX_TOL
    h = np.load(os.path.join(golden_dir, "ref_helpers.npz"))
    got = np.array([wb.calculate_far_field_delays(a, 0.04, 343.0) for a in h["ffd_angles"]])
    assert np.array_equal(got, h["ffd_out"])
    # zero delay is the identity up to float32 transform noise; two delays compose like the reference's (not exactly
    # additive: each irfft drops the imaginary part the ramp gave the Nyquist bin)
    y = torch.from_numpy(z["fd2_in"]).cuda()
    assert rel_l2(wb.apply_frac_delay(y, 0.0, 16000).cpu().numpy(), z["fd2_in"]) < MIX_TOL
    two = wb.apply_frac_delay(wb.apply_frac_delay(y, 2e-5, 16000), 1.5e-5, 16000)
    ref2 = O.fractional_delay(O.fractional_delay(z["fd2_in"].astype(np.float64), 2e-5, 16000), 1.5e-5, 16000)
    assert rel_l2(two.cpu().numpy(), ref2) < 2 * MIX_TOL


# lengths with no 2^a * (<= 1024) split run the chirp-z (Bluestein) path: primes, odd lengths, 2 * prime, tiny
@pytest.mark.parametrize("B,S,L", [(1, 3, 4001), (2, 4, 12345), (1, 2, 50001), (2, 1, 6002), (1, 5, 2049), (1, 8, 100003),
                                     (1, 3, 262143)])
def test_mixer_arbitrary_length_matches_oracle(az, B, S, L):
    rng = np.random.default_rng(7000 * S + L % 991)
    src = rng.standard_normal((B, S, L)).astype(np.float32)
    delays = [O.far_field_delays(a, 0.04, 343.0) for a in ANGLES[:S]]
    assert az._lib.load().avz_farfield_mix_ws_bytes(B, S, L) > B * 2 * L * 8 * 2   # chirp-z scratch, not the native one
    mix, tgt, itf = az.ops.far_field_mix(torch.from_numpy(src).cuda(), delays, 16000.0)
    assert mix.shape == (B, 2, L) and tgt.shape == (B, L) and itf.shape == (B, L)
    for b in range(B):
        rm, rt, ri = O.mix_far_field(list(src[b].astype(np.float64)), ANGLES[:S], 0.04, 343.0, 16000.0)
        assert rel_l2(mix[b].cpu().numpy(), rm) < MIX_TOL
        assert rel_l2(tgt[b].cpu().numpy(), rt) < MIX_TOL
        if S > 1:
            assert rel_l2(itf[b].cpu().numpy(), ri) < MIX_TOL
        assert abs(float(mix[b].abs().max()) - 1.0) < 1e-6
    # the cached plan serves a second call (other stream, other batch) identically
    with torch.cuda.stream(torch.cuda.Stream()):
        mix2, _, _ = az.ops.far_field_mix(torch.from_numpy(src[:1]).cuda(), delays, 16000.0)
    torch.cuda.synchronize()
    assert torch.equal(mix2[0], mix[0])


def test_apply_frac_delay_arbitrary_length(az):
    from avzoom.core import world_building as wb
    rng = np.random.default_rng(11)
    for L in (33333, 16001):
        y = rng.standard_normal(L).astype(np.float32)
        got = wb.apply_frac_delay(y, 3.7e-5, 16000)
        assert isinstance(got, np.ndarray) and got.shape == (L,)
        assert rel_l2(got, O.fractional_delay(y.astype(np.float64), 3.7e-5, 16000)) < MIX_TOL


def test_mixer_rejects_unsupported_shapes(az):
    x = torch.zeros((1, 3, 300001), device="cuda")          # prime and 2L-1 > 512 * 1024: neither path takes it
    with pytest.raises(az._lib.AvzError):
        az.ops.far_field_mix(x, [[0.0, 0.0]] * 3)
    x = torch.zeros((1, 9, 4000), device="cuda")            # more than 8 sources
    with pytest.raises(az._lib.AvzError):
        az.ops.far_field_mix(x, [[0.0, 0.0]] * 9)


def test_mixer_feeds_the_hot_path(az):
    """Mixtures made on the device go straight into the fused oracle-mask MVDR: same result as host-mixed input."""
    rng = np.random.default_rng(5)
    L = 16000
    from avzoom import synth
    src = np.stack([np.stack([synth.speech_like(rng, L) for _ in range(3)]) for _ in range(2)]).astype(np.float32)
    delays = [O.far_field_delays(a, 0.04, 343.0) for a in ANGLES[:3]]
    mix, tgt, itf = az.ops.far_field_mix(torch.from_numpy(src).cuda(), delays)
    out = az.ops.oracle_mask_mvdr(mix, tgt, itf)
    sc = az.ops.sir_scores(out, tgt, itf).cpu().numpy()
    sc_in = az.ops.sir_scores(mix[:, 0].contiguous(), tgt, itf).cpu().numpy()
    assert np.all(sc[:, 1] - sc_in[:, 1] > 10.0)            # the beamformer improves OSIR by > 10 dB
    for b in range(2):
        ref = O.oracle_mask_mvdr(mix[b].cpu().numpy().astype(np.float64), tgt[b].cpu().numpy().astype(np.float64),
                                 itf[b].cpu().numpy().astype(np.float64))
        assert rel_l2(out[b].cpu().numpy(), ref) < 1e-4


@pytest.mark.parametrize("n", [1, 7, 8, 4099, 1 << 20])
def test_pcm16_wire_format(az, n):
    rng = np.random.default_rng(n)
    pcm = rng.integers(-32768, 32768, size=n, dtype=np.int64).astype(np.int16)
    pcm[:3] = [-32768, 32767, 0][:min(3, n)]
    f = az.ops.pcm16_to_float(torch.from_numpy(pcm).cuda())
    assert np.array_equal(f.cpu().numpy(), pcm.astype(np.float32) / 32768.0)
    x = (rng.standard_normal(n) * 0.5).astype(np.float32)
    x[:min(5, n)] = np.array([1.5, -1.5, 0.5 / 32767.0, 1.5 / 32767.0, np.nan], dtype=np.float32)[:min(5, n)]
    want = np.clip(np.rint(np.nan_to_num(x.astype(np.float32) * np.float32(32767.0), nan=0.0)), -32768, 32767).astype(np.int16)
    got = az.ops.float_to_pcm16(torch.from_numpy(x).cuda()).cpu().numpy()
    assert np.array_equal(got, want)
    # what soundfile does: write then read is idempotent after the first quantisation
    again = az.ops.float_to_pcm16(az.ops.pcm16_to_float(torch.from_numpy(got).cuda()) * (32768.0 / 32767.0))
    assert np.abs(again.cpu().numpy().astype(np.int32) - got.astype(np.int32)).max() <= 1


def test_host_pipeline_pcm16_wire(az):
    """HostPipeline(wire='pcm16'): int16 host buffers in and out equal the float path run on pcm/32768 and quantised
    like soundfile.write does."""
    from avzoom import pipeline, synth
    cfg = az.PRESETS["baseline_oracle"]
    dev = torch.device("cuda", 0)
    mix, tgt, itf = synth.make_batch(2, 8, 1.0, 3)
    q = lambda a: np.clip(np.rint(a * 32767.0), -32768, 32767).astype(np.int16)
    mw, tw, iw = q(mix), q(tgt), q(itf)
    eng = pipeline.OracleMvdr(cfg, 8, 16000, dev)
    hp = pipeline.HostPipeline(eng, 1, sub_batches=4, wire="pcm16")
    pin = lambda a: torch.from_numpy(a).pin_memory()
    out, sc = hp.run(pin(mw), pin(tw), pin(iw))
    assert out.dtype == torch.int16 and out.shape == (8, eng.out_len)
    f = lambda a: torch.from_numpy(a.astype(np.float32) / 32768.0).cuda()
    ref = az.ops.oracle_mask_mvdr(f(mw), f(tw), f(iw))
    want = az.ops.float_to_pcm16(ref).cpu().numpy().astype(np.int32)
    assert np.abs(out.numpy().astype(np.int32) - want).max() <= 1
    ref_sc = az.ops.sir_scores(ref, f(tw), f(iw)).cpu().numpy()
    assert np.allclose(sc.numpy(), ref_sc, atol=1e-3)
    with pytest.raises(az._lib.AvzError):
        hp.run(pin(mix), pin(tgt), pin(itf))                # float buffers on the int16 wire


@pytest.mark.parametrize("B,S,L", [(5, 4, 64000), (2, 3, 80000), (3, 1, 16000), (2, 2, 4096), (1, 4, 96), (3, 3, 840),
                                     (40, 4, 16000), (1, 4, 8), (2, 4, 96000)])
def test_cluster_resident_mixer_against_oracle_and_multi_pass(az, B, S, L):
    """The one-kernel mixer (8-CTA cluster per utterance, spectra on chip) and the multi-pass one are two float32
    factorisations of the same transform: both within MIX_TOL of the float64 oracle, and of each other; more utterances
    than clusters (B = 40) exercises the persistent loop; reruns are bit-identical."""
    rng = np.random.default_rng(7 * S + L % 991)
    src = rng.standard_normal((B, S, L)).astype(np.float32)
    src[:, :, : L // 3] *= 0.05
    delays = [O.far_field_delays(a, 0.04, 343.0) for a in ANGLES[:S]]
    x = torch.from_numpy(src).cuda()
    one = az.ops.far_field_mix(x, delays, 16000.0)
    many = az.ops.far_field_mix(x, delays, 16000.0, multi_pass=True)
    again = az.ops.far_field_mix(x, delays, 16000.0)
    for u, v, w in zip(one, many, again):
        assert torch.equal(u, w)
        assert rel_l2(u.cpu().numpy(), v.cpu().numpy()) < MIX_TOL or float(v.abs().max()) < 1e-6
    for b in sorted({0, B // 2, B - 1}):
        rm, rt, ri = O.mix_far_field(list(src[b].astype(np.float64)), ANGLES[:S], 0.04, 343.0, 16000.0)
        assert rel_l2(one[0][b].cpu().numpy(), rm) < MIX_TOL
        assert rel_l2(one[1][b].cpu().numpy(), rt) < MIX_TOL
        if S > 1:
            assert rel_l2(one[2][b].cpu().numpy(), ri) < MIX_TOL
        else:
            assert float(one[2][b].abs().max()) < 1e-6
        assert abs(float(one[0][b].abs().max()) - 1.0) < 1e-6
    # no normalisation: plain delays (apply_frac_delay's use of the operator)
    raw = az.ops.far_field_mix(x[:1], delays, 16000.0, peak_eps=None)
    raw2 = az.ops.far_field_mix(x[:1], delays, 16000.0, peak_eps=None, multi_pass=True)
    assert rel_l2(raw[0].cpu().numpy(), raw2[0].cpu().numpy()) < MIX_TOL

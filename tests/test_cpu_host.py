"""CPU-side checks: the C-ABI library loads and exports every symbol include/avzoom.h declares, host-only
entry points agree with scipy, configuration presets, the synthetic mixer.  No GPU needed."""
import ctypes
import os
import re

import numpy as np
import pytest

import oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def az():
    import __graft_entry__ as ge
    ge.build()
    import avzoom
    return avzoom


def test_library_exports_every_declared_symbol(az):
    hdr = open(os.path.join(ROOT, "include", "avzoom.h")).read()
    declared = set(re.findall(r"AVZ_API\s+[\w\s\*]+?\b(avz_\w+)\s*\(", hdr))
    assert len(declared) >= 20
    lib = ctypes.CDLL(az._lib.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in avzoom.h but not exported"
    assert declared == set(az._lib.SIGNATURES), "ctypes table and header disagree"
    assert az._lib.load().avz_version() == 100
    info = az._lib.load().avz_build_info().decode()
    assert "sm_100a" in info and "release" in info      # the in-tree library is the release build


def test_cfg_struct_matches_header(az):
    hdr = open(os.path.join(ROOT, "include", "avzoom.h")).read()
    body = re.search(r"typedef struct AvzMvdrCfg \{(.*?)\} AvzMvdrCfg;", hdr, re.S).group(1)
    fields = re.findall(r"^\s*(float|int32_t)\s+(\w+);", body, re.M)
    assert [n for _, n in fields] == [n for n, _ in az._lib.AvzMvdrCfg._fields_]
    assert ctypes.sizeof(az._lib.AvzMvdrCfg) == 4 * len(fields)


@pytest.mark.parametrize("L", [512, 513, 639, 640, 641, 1000, 21920, 32000, 64000, 64001, 80000, 117143])
@pytest.mark.parametrize("n_fft,hop", [(512, 128), (512, 256), (1024, 512), (256, 64), (1024, 128)])
def test_num_frames_matches_scipy(az, L, n_fft, hop):
    if L < n_fft:
        pytest.skip("shorter than a frame")
    assert az.num_frames(L, n_fft, hop) == O.n_frames(L, n_fft, hop)
    assert az.num_frames(L, n_fft, hop) == -(-L // hop) + 1


def test_presets_match_oracle_presets(az):
    for name in ("baseline_oracle", "oracle_debug", "masked_mvdr", "full_audio", "tf_lite"):
        a, b = az.PRESETS[name], O.PRESETS[name]
        for f in ("fs", "n_fft", "hop", "mic_dist", "c", "angle_deg", "sigma", "hp_hz", "hp_mode", "sqrt_eps", "norm_eps",
                  "w_eps", "post", "post_floor", "peak_eps"):
            assert getattr(a, f) == getattr(b, f), (name, f)
    assert az.PRESETS["baseline_oracle"].hp_bins() == 4        # f < 100 Hz at 512 / 16 kHz: k = 0..3
    assert az.PRESETS["full_audio"].hp_bins() == 7             # k = 0..6 at 1024
    assert az.PRESETS["tf_lite"].to_c().hp_mode == az._lib.HP_NONE


def test_no_cpu_fallback(az):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(az._lib.AvzError):
        az.stft(np.zeros((2, 4000), np.float32), 512, 128)


def test_synthetic_mixer_follows_world_building(az):
    from avzoom import synth
    mix, tgt, itf = synth.make_mixture(1_000_003 * 2 + 5, 16000, 3)
    assert mix.shape == (2, 16000) and mix.dtype == np.float32
    assert abs(np.abs(mix).max() - 1.0) < 1e-6
    # mic-1 mixture is the sum of the two references (world_building.py:77-85)
    assert np.max(np.abs(mix[0] - (tgt + itf))) < 1e-6
    # the same sources through the oracle's restatement of mix_and_save
    rng = np.random.default_rng(1_000_003 * 2 + 5)
    srcs = [synth.speech_like(rng, 16000) for _ in range(4)]
    m2, t2, i2 = O.mix_far_field(srcs, [90.0, 40.0, 130.0, 65.0])
    assert np.max(np.abs(m2 - mix)) < 1e-6 and np.max(np.abs(t2 - tgt)) < 1e-6 and np.max(np.abs(i2 - itf)) < 1e-6
    # the IBM is non-trivial on this material
    frac = O.ibm_noise_mask(O.stft_scipy(tgt, 512, 128), O.stft_scipy(itf, 512, 128)).mean()
    assert 0.3 < frac < 0.9
    # deterministic and batchable
    b1 = synth.make_batch(2, 3, 1.0, 3, start=5)
    assert np.array_equal(b1[0][0], mix[:, :16000])


def test_release_library_reads_no_environment_variable():
    """Every getenv in csrc/ sits behind #ifdef AVZ_EXPERIMENT (tools/build_exp.sh): the results of the release library
    depend on its arguments only (a stray AVZ_IBM_TOL must not change IBM bits)."""
    csrc = os.path.join(ROOT, "real-time-audio-visual-zooming_b200", "csrc")
    for name in sorted(os.listdir(csrc)):
        if not name.endswith((".cu", ".cuh")):
            continue
        depth_exp, stack = 0, []
        for ln, line in enumerate(open(os.path.join(csrc, name)), 1):
            t = line.strip()
            if t.startswith("#if"):
                stack.append("AVZ_EXPERIMENT" in t and not t.startswith("#ifndef"))
            elif t.startswith("#else") and stack:
                stack[-1] = False
            elif t.startswith("#endif") and stack:
                stack.pop()
            if "getenv(" in t and not t.startswith("//"):
                assert any(stack), f"{name}:{ln}: getenv outside #ifdef AVZ_EXPERIMENT"

"""CPU-side checks of the reference-mirroring modules: steering vectors against the reference's own outputs
(tests/golden), mask-model definitions (state-dict names and the seeded random-init output the reference's class
produces), WAV I/O."""
import os

import numpy as np
import pytest
import torch


@pytest.fixture(scope="module")
def helpers(golden_dir):
    return np.load(os.path.join(golden_dir, "ref_helpers.npz"))


def test_steering_vector_dropins_bit_exact(helpers):
    from avzoom.core import masked_mvdr, batch_mvdr, chunked
    from avzoom.final_pipeline import inference as fpi
    for args, ref in zip(helpers["sv_args"], helpers["sv_out"]):
        got = masked_mvdr.get_steering_vector(*args)
        assert got.shape == (2, 1) and got.dtype == np.complex128
        assert np.array_equal(got[:, 0], ref)
        assert np.array_equal(chunked.get_steering_vector(*args)[:, 0], ref)
    f = helpers["asv_f_bins"]
    assert np.array_equal(batch_mvdr.get_all_steering_vectors(f, 90.0, 0.04, 343.0), helpers["asv_out"])
    assert np.array_equal(batch_mvdr.get_all_steering_vectors(f, 40.0, 0.04, 343.0), helpers["asv_out_40"])
    v = fpi.get_steering_vector_single(1000.0, 90.0, 0.08, 343.0)
    assert v.shape == (2, 1) and abs(v[0, 0] - 1.0) < 1e-9
    consts = [masked_mvdr.FS, masked_mvdr.D, masked_mvdr.C, masked_mvdr.ANGLE_TARGET, masked_mvdr.N_MICS,
              masked_mvdr.SIGMA, masked_mvdr.N_FFT, masked_mvdr.N_HOP]
    assert consts == helpers["constants_masked_mvdr"].tolist()


def test_mask_models_match_reference_definition(golden_dir):
    from avzoom.core import models
    import oracle as O
    lc = np.load(os.path.join(golden_dir, "ref_learned_chunk.npz"))
    sp = np.load(os.path.join(golden_dir, "ref_speech_excerpt.npz"))
    torch.manual_seed(0)
    m = models.FreqPreservingUNet().eval()
    assert list(m.state_dict().keys()) == list(lc["state_keys"])
    assert sum(p.numel() for p in m.parameters()) == 1842113
    assert sum(p.numel() for p in models.DeepFPU().parameters()) == 16051009
    # same seed, same construction order -> same random weights -> the mask the reference's class produced
    mix = (sp["mix_pcm"].astype(np.float32) / 32768.0)[:32000]
    import scipy.signal
    Y = scipy.signal.stft(mix.T, fs=16000, nperseg=1024, noverlap=512)[2]
    X = torch.from_numpy(np.stack([np.log(np.abs(Y[0]) + 1e-7), np.angle(Y[0]) - np.angle(Y[1])], 0)).float()[None]
    with torch.no_grad():
        mask = m(X)[0].numpy()
    assert mask.shape == (513, 64)
    assert np.max(np.abs(mask - lc["masks"][0])) < 1e-5
    d = models.DeepFPU().eval()
    with torch.no_grad():
        out = d(X[:, :, :, :33])            # odd time length exercises the nearest-neighbour resize
    assert out.shape == (1, 513, 33) and float(out.min()) > 0 and float(out.max()) < 1


def test_wav_roundtrip(tmp_path):
    from avzoom import wavio
    rng = np.random.default_rng(0)
    pcm = rng.integers(-32768, 32767, size=(1000, 2)).astype(np.int16)
    import wave
    p = str(tmp_path / "a.wav")
    with wave.open(p, "wb") as w:
        w.setnchannels(2); w.setsampwidth(2); w.setframerate(16000); w.writeframes(pcm.tobytes())
    x, fs = wavio.read(p)
    assert fs == 16000 and x.dtype == np.float32 and x.shape == (1000, 2)
    assert np.array_equal(x, pcm.astype(np.float32) / 32768.0)
    q = str(tmp_path / "b.wav")
    wavio.write(q, np.array([0.0, 0.5, -0.5, 1.0, -1.0, 2.0]), 16000)
    y, _ = wavio.read(q)
    assert np.array_equal(np.rint(y * 32768).astype(int), [0, 16384, -16384, 32767, -32767, 32767])


def test_chunk_split_matches_reference_windows():
    from avzoom.core import chunked
    L, win = 40000, 32000
    y = torch.arange(L * 2, dtype=torch.float32).reshape(L, 2)
    chunks, stride = chunked.split_chunks(y, win)
    assert stride == 16000 and chunks.shape == (3, 2, win)      # ceil(40000 / 16000) windows
    for i in range(3):
        ref = y[i * stride:i * stride + win]
        ref = torch.nn.functional.pad(ref, (0, 0, 0, win - ref.shape[0]))
        assert torch.equal(chunks[i], ref.T)
    assert chunked.n_windows(L, win) == 3 and chunked.n_windows(32000, win) == 2 and chunked.n_windows(16001, win) == 2
    # the overlap-add is a device kernel (avz_chunk_ola_f32): no CPU path, it must fail loudly without a GPU
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            chunked.overlap_add_chunks(torch.ones((3, 32256)), L, win, stride, buf_extra=win)


def test_tflite_beamformer_wrappers_keep_the_reference_contract(tmp_path):
    """TFLiteBeamformer.predict_mask (Final_pipeline/src/inference.py:117-141, tf_lite_version/inference.py:205-241):
    the model sees the reference's NHWC input tensor and the result squeezes to (F, T)."""
    from avzoom.final_pipeline import inference as fin
    from avzoom.core import batch_mvdr as tfl
    rng = np.random.default_rng(3)
    F, T = fin.FREQ_BINS, 7
    lm = rng.standard_normal((F, T)).astype(np.float32)
    ipd = rng.uniform(-6, 6, (F, T)).astype(np.float32)
    seen = {}

    def model4(x):
        seen["x4"] = x.clone()
        return torch.sigmoid(x[..., 0] + x[..., 3])[..., None]

    out = fin.TFLiteBeamformer("unused", model=model4).predict_mask(lm, ipd)
    want = np.stack([lm, np.sin(ipd), np.cos(ipd), np.tile(np.linspace(0, 1, F, dtype=np.float32)[:, None], (1, T))], axis=-1)[None]
    assert isinstance(out, np.ndarray) and out.shape == (F, T)
    assert seen["x4"].shape == (1, F, T, 4) and np.allclose(seen["x4"].numpy(), want, atol=1e-6)
    with pytest.raises(FileNotFoundError):
        fin.TFLiteBeamformer(str(tmp_path / "missing.tflite"))

    def model2(x):
        seen["x2"] = x.clone()
        return x[..., 1]

    out2 = tfl.TFLiteBeamformer(model=model2).predict_mask(lm, ipd)
    assert seen["x2"].shape == (1, F, T, 2) and np.array_equal(out2, ipd)
    # a TorchScript file stands in for the .tflite blob
    class M(torch.nn.Module):
        def forward(self, x):
            return torch.sigmoid(x[..., 0])
    path = str(tmp_path / "mask.pt")
    torch.jit.script(M()).save(path)
    bf = tfl.TFLiteBeamformer(path)
    x = torch.from_numpy(lm).to(next(iter([torch.device("cuda" if torch.cuda.is_available() else "cpu")])))
    got = bf.predict_mask(x, torch.from_numpy(ipd).to(x.device))
    assert torch.allclose(got.cpu(), torch.sigmoid(torch.from_numpy(lm)), atol=1e-6)

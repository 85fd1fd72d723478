"""GPU tests of the chunk drivers: windows read in place, count-averaged
overlap-add as a kernel, batches of recordings - against the reference's own outputs (tests/golden) and the oracle."""
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
    return avzoom


def rel_l2(a, b):
    a = np.asarray(a, dtype=np.float64)
    return float(np.linalg.norm(a - np.asarray(b)) / (np.linalg.norm(b) + 1e-300))


class Replay(torch.nn.Module):
    """Stands in for the mask model: returns pre-computed masks (one per window)."""

    def __init__(self, masks):
        super().__init__()
        self.masks = torch.from_numpy(np.ascontiguousarray(masks)).cuda()

    def forward(self, X):
        return self.masks[:X.shape[0]]


class CheapMask(torch.nn.Module):
    """A deterministic pointwise 'mask model' (the U-Net is not what these tests are about)."""

    def forward(self, X):
        return torch.sigmoid(0.5 * X[:, 0] + 3.0 + 0.3 * torch.cos(X[:, 1]))


def _write_wav(path, pcm):
    pcm = np.asarray(pcm, dtype="<i2")
    with wave.open(path, "wb") as w:
        w.setnchannels(1 if pcm.ndim == 1 else pcm.shape[1])
        w.setsampwidth(2)
        w.setframerate(16000)
        w.writeframes(pcm.tobytes())


def test_chunk_ola_kernel_matches_reference_loops(az):
    """avz_chunk_ola_f32 against the two buffer conventions of the reference written as their own loops."""
    from avzoom.core import chunked
    rng = np.random.default_rng(0)
    win, stride, olen = 32000, 16000, 32256
    for L in (40000, 32000, 16001, 100000):
        n = chunked.n_windows(L, win)
        outs = rng.standard_normal((n, olen)).astype(np.float32)
        # main_deploy: buffers of L + WIN, min(len, WIN) samples per window (full_audio.../inference.py:135-156)
        ob, nb = np.zeros(L + win, np.float32), np.zeros(L + win, np.float32)
        for i in range(n):
            k = min(olen, win)
            ob[i * stride:i * stride + k] += outs[i, :k]
            nb[i * stride:i * stride + k] += 1
        nb[nb == 0] = 1
        ref = (ob / nb)[:L]
        got = chunked.overlap_add_chunks(torch.from_numpy(outs).cuda(), L, win, stride, buf_extra=win).cpu().numpy()
        assert np.array_equal(got, ref)
        # TFLite-era: buffers of L, all iSTFT samples up to the end of the buffer (Final_pipeline/src/inference.py:174-233);
        # 256 samples are covered by three windows
        ob, nb = np.zeros(L, np.float32), np.zeros(L, np.float32)
        for i in range(n):
            k = min(olen, L - i * stride)
            ob[i * stride:i * stride + k] += outs[i, :k]
            nb[i * stride:i * stride + k] += 1
        ref = ob / np.maximum(nb, 1)
        got = chunked.overlap_add_chunks(torch.from_numpy(outs).cuda(), L, win, stride, buf_extra=0).cpu().numpy()
        assert np.array_equal(got, ref)
        if L >= 48256:
            assert nb.max() == 3


def test_pcm16_frames_to_planar(az):
    from avzoom.core import chunked
    rng = np.random.default_rng(1)
    pcm = rng.integers(-32768, 32767, size=(12345, 2)).astype(np.int16)
    got = chunked.to_planar(pcm)
    assert got.shape == (1, 2, 12345) and got.dtype == torch.float32
    assert np.array_equal(got[0].cpu().numpy(), pcm.T.astype(np.float32) / 32768.0)
    f = chunked.to_planar(pcm.astype(np.float32) / 32768.0)
    assert torch.equal(f, got)


def test_windows_in_place_equal_explicit_windows(az):
    """The in-place window view (no gathered copy) gives what the explicit batch of zero-padded windows gives."""
    from avzoom.core import chunked
    from avzoom import synth
    for L in (40000, 40001, 16000, 70007):              # odd lengths: the second channel plane is not 8-byte aligned
        mix, _, _ = synth.make_batch(9, 1, L / 16000.0, 2)
        rec = torch.from_numpy(mix[:, :, :L].copy()).cuda()
        L = rec.shape[-1]
        cfg = az.PRESETS["full_audio"]
        model = CheapMask()
        got = chunked.ChunkedEnhancer(cfg)(rec, model)[0]
        chunks, stride = chunked.split_chunks(rec[0].T.contiguous(), 32000)
        outs = chunked.enhance_chunks(chunks, model, cfg)
        want = chunked.overlap_add_chunks(outs, L, 32000, stride, buf_extra=32000)
        assert got.shape == want.shape == (L,)
        assert rel_l2(got.cpu().numpy(), want.cpu().numpy()) < 1e-6
        X_view = torch.empty((chunks.shape[0], 2, 513, 64), dtype=torch.float32, device="cuda")
        import ctypes as C
        lib = az._lib.load()
        cv = az._lib.AvzChunkView(L, chunks.shape[0], 16000)
        az._lib.check(lib.avz_chunk_features_f32(az.ops._ptr(rec), 1, C.byref(cv), 32000, 1024, 512, 0, az.ops._ptr(X_view),
                                                 az.ops._stream()), "avz_chunk_features_f32")
        assert torch.equal(X_view, az.wave_features(chunks, 1024, 512))


def test_main_deploy_and_tflite_era_drivers_match_reference_output(az, golden_dir, tmp_path):
    """main_deploy (full_audio.../inference.py:120-167), process_audio_file (tf_lite_version/inference.py:245-391) and
    enhance_audio (Final_pipeline/src/inference.py:144-238) against what the reference's own code wrote for the same
    recording and the same masks (float64 reads: the reference's arithmetic without its complex64 STFT noise)."""
    from avzoom.core import chunked, batch_mvdr
    from avzoom.final_pipeline import inference as fpi, config as fcfg
    from avzoom import wavio
    g = np.load(os.path.join(golden_dir, "ref_speech_excerpt.npz"))
    lc = np.load(os.path.join(golden_dir, "ref_learned_chunk.npz"))
    d5 = np.load(os.path.join(golden_dir, "ref_chunk_drivers.npz"))
    L = int(lc["L"])
    pcm = g["mix_pcm"][:L]
    model = Replay(lc["masks"])
    full = chunked.enhance_waveform(pcm.astype(np.float32) / 32768.0, model, az.PRESETS["full_audio"])
    assert rel_l2(full, lc["main_deploy_out"]) < 1e-4
    full_pcm = chunked.enhance_waveform(pcm, model, az.PRESETS["full_audio"])        # raw PCM16 frames, converted on the device
    assert np.array_equal(full_pcm, full)
    wav = str(tmp_path / "speech_TEST.wav")
    _write_wav(wav, pcm)
    out = batch_mvdr.process_audio_file(wav, str(tmp_path / "paf.wav"), model=model)
    assert out.shape == (L,)
    assert rel_l2(out, d5["process_audio_file_out_f64read"]) < 1e-4
    assert abs(np.max(np.abs(out)) - 1.0) < 1e-6                                      # final / (max + 1e-9)
    # enhance_audio: hybrid hard-null beamformer, post-filter S * mask, written as PCM16
    old = fcfg.RESULTS_DIR
    fcfg.RESULTS_DIR = str(tmp_path)
    try:
        p = fpi.enhance_audio("g5", wav, None, model=model)
    finally:
        fcfg.RESULTS_DIR = old
    ref = d5["enhance_audio_out_f64read"]
    enh = chunked.ChunkedEnhancer(fpi.FINAL_CFG, fcfg.WIN_SIZE, clip_to_input=True, weights="hybrid_null", final_peak_eps=1e-9,
                                  steering=lambda dev: fpi._steering(fpi.FINAL_CFG.freqs(), dev, wide=True))
    x = enh(chunked.to_planar(pcm), model)[0].cpu().numpy()
    assert rel_l2(x, ref) < 1e-4
    y, _ = wavio.read(p, dtype="float64")
    q = (1.0 / 32767.0) / np.sqrt(12.0) * np.sqrt(len(ref)) / np.linalg.norm(ref)
    assert rel_l2(y * (32768.0 / 32767.0), ref) < 1.5 * q + 1e-4


def test_batch_of_256_recordings(az):
    """256 recordings x 4 s in one call (B = 1024 windows): every recording equals its own single-recording run to
    float32 summation-order noise, and two of them are checked against the float64 oracle's main_deploy restatement."""
    from avzoom.core import chunked
    from avzoom import synth
    mix, _, _ = synth.make_batch(7, 8, 4.0, 2)
    rec = torch.from_numpy(np.tile(mix, (32, 1, 1))).cuda()
    rec[8:] *= torch.linspace(0.5, 1.5, 248, device="cuda")[:, None, None]         # 256 distinct recordings
    cfg = az.PRESETS["full_audio"]
    model = CheapMask()
    enh = chunked.ChunkedEnhancer(cfg)
    out = enh(rec, model, timing=True)
    assert out.shape == (256, 64000) and bool(torch.isfinite(out).all())
    assert set(enh.last_ms) == {"features", "mask_model", "mvdr_and_chunk_ola"}
    for r in (0, 7, 100, 255):
        single = enh(rec[r:r + 1].contiguous(), model)[0]
        assert rel_l2(out[r].cpu().numpy(), single.cpu().numpy()) < 1e-5

    def mask_fn(X):
        return CheapMask()(torch.from_numpy(X[None].astype(np.float32))).numpy()[0]

    for r in (3, 200):
        y = rec[r].T.cpu().numpy().astype(np.float64)
        ref = O.chunked_enhance(y, mask_fn, O.PRESETS["full_audio"], win=32000)
        assert rel_l2(out[r].cpu().numpy(), ref) < 1e-4

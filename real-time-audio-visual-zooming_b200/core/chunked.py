"""Learned-mask chunk path: drop-in for `process_chunk` / `main_deploy` of
rt_av_zoom/core/full_audio_generating_pipeline/inference.py (:88-167) and rt_av_zoom/core/resnet_model_mvdr/inference.py
(:152-275).  The sliding 2 s windows of one recording become the batch dimension of the fused kernels:
features (STFT fused) -> mask model -> masked covariance -> weights -> beamform + post-filter + iSTFT, then the
count-averaged overlap-add of the chunks."""
from __future__ import annotations

import os
import time

import numpy as np
import torch

from .. import ops, wavio
from ..config import MvdrConfig, PRESETS
from .models import DeepFPU, FreqPreservingUNet, ResBlock  # noqa: F401

CONF = {"fs": 16000, "n_fft": 1024, "hop_len": 512, "d": 0.04, "c": 343.0, "train_seg_samples": 32000}  # config.json
FS, N_FFT, HOP, WIN_SIZE_SAMPLES, D, C = (CONF["fs"], CONF["n_fft"], CONF["hop_len"], CONF["train_seg_samples"],
                                          CONF["d"], CONF["c"])
ANGLE_TARGET = 90.0
SIGMA = 1e-5
N_MICS = 2


def get_steering_vector(angle_deg, f, d, c):
    """full_audio.../inference.py:70-75."""
    from .masked_mvdr import get_steering_vector as _sv
    return _sv(angle_deg, f, d, c)


def calculate_metrics_manual(output, target, interf):
    """full_audio.../inference.py:77-85: returns the projection SIR twice."""
    if len(target) == 0:
        return 0, 0
    sc = ops.sir_scores(np.asarray(output, np.float32), np.asarray(target, np.float32), np.asarray(interf, np.float32))
    return float(sc[3]), float(sc[3])


def load_mask_model(path=None, arch="unet"):
    model = FreqPreservingUNet() if arch == "unet" else DeepFPU()
    if path is not None:
        model.load_state_dict(torch.load(path, map_location="cpu"))
    return model.eval().cuda()


def n_windows(rec_len: int, win: int) -> int:
    """ceil(L / (win / 2)) windows (full_audio.../inference.py:137)."""
    return int(np.ceil(rec_len / (win // 2)))


def split_chunks(y_full: torch.Tensor, win: int):
    """(L, 2) -> chunks (n, 2, win): the reference's windows as an explicit (copied) batch, for callers that want them
    materialised.  The enhancement path below does NOT use this: its kernels read the windows in place."""
    L = y_full.shape[0]
    stride = win // 2
    n = n_windows(L, win)
    padded = torch.zeros((n * stride + win, y_full.shape[1]), dtype=y_full.dtype, device=y_full.device)
    padded[:L] = y_full
    idx = torch.arange(n, device=y_full.device)[:, None] * stride + torch.arange(win, device=y_full.device)[None, :]
    return padded[idx].permute(0, 2, 1).contiguous(), stride


def enhance_chunks(chunks: torch.Tensor, model, cfg: MvdrConfig) -> torch.Tensor:
    """process_chunk for an explicit batch of windows: chunks (n, 2, win) -> (n, iSTFT length)."""
    X = ops.wave_features(chunks, cfg.n_fft, cfg.hop, "logmag_ipd")
    with torch.no_grad():
        mask = model(X).float().contiguous()
    return ops.learned_mask_mvdr(chunks, mask, cfg)


class ChunkedEnhancer:
    """The chunk drivers of the reference as one batched device pipeline:

        rec [R, 2, Lrec] planar float32  ->  features of every 2 s window (read in place, no gathered copy)
          -> mask model  ->  masked covariance  ->  weights  ->  beamform + post-filter + iSTFT per window
          -> count-averaged overlap-add of the window outputs  ->  final [R, Lrec]   (-> optional peak normalisation)

    Between the waveform and `final` the only torch operation is the mask model; everything else is one C-ABI call per
    stage (avz_chunk_features_f32, avz_chunk_mask_cov_f32, avz_mvdr_weights_f32 / avz_hybrid_null_weights_f64,
    avz_chunk_mvdr_apply_f32, avz_chunk_ola_f32).  `clip_to_input` = False is main_deploy (buffer of L + WIN, a window
    contributes min(len, WIN) samples: full_audio.../inference.py:135-153); True is the TFLite-era drivers (buffer of L,
    a window contributes all its iSTFT samples up to the end of the buffer: tf_lite_version/inference.py:267-373,
    Final_pipeline/src/inference.py:174-227).  `weights`: 'mvdr' or 'hybrid_null' (Final_pipeline/src/inference.py:28-98,
    covariance and solve in float64)."""

    def __init__(self, cfg: MvdrConfig, win: int = WIN_SIZE_SAMPLES, clip_to_input: bool = False, weights: str = "mvdr",
                 features: str = "logmag_ipd", final_peak_eps: float | None = None, steering=None):
        from .. import _lib
        if cfg.n_fft != 1024 or cfg.hop != 512:
            raise _lib.AvzError("the chunk drivers run at n_fft 1024 / hop 512 (the reference's config.json)")
        if cfg.peak_eps is not None:
            raise ValueError("per-window peak normalisation is not part of any chunk driver: use final_peak_eps")
        self.cfg, self.win, self.stride = cfg, int(win), int(win) // 2
        self.clip, self.weights, self.features, self.final_peak_eps = clip_to_input, weights, features, final_peak_eps
        self.lib = _lib.load()
        self._lib = _lib
        self.T = ops.num_frames(self.win, cfg.n_fft, cfg.hop)
        self.olen = (self.T - 1) * cfg.hop
        self.use_len = self.olen if clip_to_input else min(self.olen, self.win)
        self.steering = steering
        self.last_ms = {}

    def __call__(self, rec: torch.Tensor, model, timing: bool = False) -> torch.Tensor:
        """rec [R, 2, Lrec] float32 CUDA (planar) -> final [R, Lrec] float32."""
        import ctypes as C
        lib, _lib, cfg = self.lib, self._lib, self.cfg
        if rec.dim() != 3 or rec.shape[1] != 2 or rec.dtype != torch.float32 or not rec.is_cuda or not rec.is_contiguous():
            raise ValueError("rec must be a contiguous float32 CUDA tensor [R, 2, Lrec]")
        R, _, Lrec = rec.shape
        n = n_windows(Lrec, self.win)
        B, F, T, dev = R * n, cfg.n_freq, self.T, rec.device
        cv = _lib.AvzChunkView(Lrec, n, self.stride)
        st, p = ops._stream, ops._ptr
        f32 = dict(dtype=torch.float32, device=dev)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)] if timing else None
        mode = {"logmag_ipd": _lib.FEAT_LOGMAG_IPD, "physics": _lib.FEAT_PHYSICS_NHWC}[self.features]
        X = torch.empty((B, F, T, 4) if mode == _lib.FEAT_PHYSICS_NHWC else (B, 2, F, T), **f32)
        if timing:
            ev[0].record()
        _lib.check(lib.avz_chunk_features_f32(p(rec), R, C.byref(cv), self.win, cfg.n_fft, cfg.hop, mode, p(X), st()),
                   "avz_chunk_features_f32")
        if timing:
            ev[1].record()
        with torch.no_grad():
            mask = model(X).float().reshape(B, F, T).contiguous()
        if timing:
            ev[2].record()
        cc = cfg.to_c()
        outs = torch.empty((B, self.olen), **f32)
        if self.weights == "hybrid_null":
            Rp = torch.empty((B, F, 4), dtype=torch.float64, device=dev)
            ms = torch.empty((B, F), dtype=torch.float64, device=dev)
            ws = torch.empty((int(lib.avz_wave_mask_cov_f64_ws_bytes(B, self.win, cfg.n_fft, cfg.hop)),), dtype=torch.uint8, device=dev)
            _lib.check(lib.avz_chunk_mask_cov_f64(p(rec), p(mask), R, C.byref(cv), self.win, cfg.n_fft, cfg.hop,
                                                  float(cfg.sqrt_eps), float(cfg.norm_eps), p(Rp), p(ms), p(ws), st()),
                       "avz_chunk_mask_cov_f64")
            w = ops.hybrid_null_weights(Rp, self.steering(dev), cfg.hp_bins(), round_to_f32=True)
            spec = None
        else:
            Rp = torch.empty((B, F, 4), **f32)
            ms = torch.empty((B, F), **f32)
            ws = torch.empty((max(int(lib.avz_ibm_cov_ws_bytes(B, self.win, cfg.n_fft, cfg.hop)), 4),), dtype=torch.uint8, device=dev)
            nspec = int(lib.avz_spec_ws_bytes(B, self.win, cfg.n_fft, cfg.hop))
            spec = torch.empty((nspec,), dtype=torch.uint8, device=dev) if nspec > 0 else None
            _lib.check(lib.avz_chunk_mask_cov_f32(p(rec), p(mask), R, C.byref(cv), self.win, cfg.n_fft, cfg.hop,
                                                  float(cfg.sqrt_eps), float(cfg.norm_eps), p(Rp), p(ms), p(ws), p(spec), st()),
                       "avz_chunk_mask_cov_f32")
            w = torch.empty((B, F, 2), dtype=torch.complex64, device=dev)
            _lib.check(lib.avz_mvdr_weights_f32(p(Rp), p(ops.steering_vectors(cfg, dev)), B, F, C.byref(cc), p(w), st()),
                       "avz_mvdr_weights_f32")
        gain_mask = mask if cfg.post in ("floor", "mask") else None
        _lib.check(lib.avz_chunk_mvdr_apply_f32(p(rec), p(spec), p(w), p(gain_mask), R, C.byref(cv), self.win, cfg.n_fft,
                                                cfg.hop, C.byref(cc), p(outs), p(None), st()), "avz_chunk_mvdr_apply_f32")
        final = torch.empty((R, Lrec), **f32)
        peak = torch.zeros((R,), **f32) if self.final_peak_eps is not None else None
        _lib.check(lib.avz_chunk_ola_f32(p(outs), R, n, self.olen, Lrec, self.stride, self.use_len, p(final), p(peak), st()),
                   "avz_chunk_ola_f32")
        if peak is not None:
            ops.peak_normalise(final, peak, self.final_peak_eps)
        if timing:
            ev[3].record()
            torch.cuda.synchronize()
            self.last_ms = {"features": ev[0].elapsed_time(ev[1]), "mask_model": ev[1].elapsed_time(ev[2]),
                            "mvdr_and_chunk_ola": ev[2].elapsed_time(ev[3])}
        return final


def overlap_add_chunks(outs: torch.Tensor, L: int, win: int, stride: int, buf_extra: int) -> torch.Tensor:
    """Count-averaged OLA of explicit window outputs (n, olen) -> (L,) through avz_chunk_ola_f32
    (full_audio.../inference.py:135-156; Final_pipeline/src/inference.py:174-175,225-233).  `buf_extra` = win for
    main_deploy (every chunk fits, min(len, WIN) samples each), 0 for the TFLite-era drivers (buffer of L)."""
    from .. import _lib
    n, olen = outs.shape
    outs = outs.contiguous()
    final = torch.empty((1, L), dtype=torch.float32, device=outs.device)
    use = min(olen, win) if buf_extra else olen
    _lib.check(_lib.load().avz_chunk_ola_f32(ops._ptr(outs), 1, n, olen, L, stride, use, ops._ptr(final), ops._ptr(None),
                                             ops._stream()), "avz_chunk_ola_f32")
    return final[0]


def to_planar(y_full) -> torch.Tensor:
    """(L, 2) frames x channels (what soundfile.read returns) -> planar [1, 2, L] float32 on the device.  int16 frames
    (raw PCM16) are converted on the device (avz_pcm16_frames_to_planar_f32: / 32768 like soundfile); float frames are
    re-laid on the host before the upload."""
    from .. import _lib
    y = np.asarray(y_full)
    if y.dtype == np.int16:
        pcm = torch.from_numpy(np.ascontiguousarray(y)).cuda()
        out = torch.empty((1, y.shape[1], y.shape[0]), dtype=torch.float32, device=pcm.device)
        _lib.check(_lib.load().avz_pcm16_frames_to_planar_f32(ops._ptr(pcm), 1, y.shape[0], y.shape[1], ops._ptr(out),
                                                              ops._stream()), "avz_pcm16_frames_to_planar_f32")
        return out
    return torch.from_numpy(np.ascontiguousarray(y.astype(np.float32, copy=False).T)).cuda()[None]


def enhance_waveform(y_full, model, cfg: MvdrConfig = PRESETS["full_audio"], win: int = WIN_SIZE_SAMPLES,
                     buf_extra: int | None = None) -> np.ndarray:
    """The body of main_deploy: (L, 2) waveform -> (L,) enhanced waveform."""
    if buf_extra is None:
        buf_extra = win
    enh = ChunkedEnhancer(cfg, win, clip_to_input=(buf_extra == 0))
    return enh(to_planar(y_full), model)[0].cpu().numpy()


def process_chunk(y_chunk, model, chunk_idx=None):
    """full_audio.../inference.py:88-118: (N, 2) chunk -> enhanced (n,).  With `chunk_idx` it is the resnet variant
    (resnet_model_mvdr/inference.py:152-211) and returns (out, t_infer, t_mvdr)."""
    cfg = PRESETS["full_audio"]
    chunk = torch.as_tensor(np.asarray(y_chunk, dtype=np.float32)).cuda().T.contiguous()[None]
    if chunk_idx is None:
        return enhance_chunks(chunk, model, cfg)[0].cpu().numpy()
    print(f"\n--- Processing chunk {chunk_idx} ---")
    X = ops.wave_features(chunk, cfg.n_fft, cfg.hop, "logmag_ipd")
    torch.cuda.synchronize()
    tic = time.time()
    with torch.no_grad():
        mask = model(X).float().contiguous()
    torch.cuda.synchronize()
    t_infer = time.time() - tic
    print(f"Mask Estimation Time: {t_infer * 1000:.2f} ms")
    tic = time.time()
    out = ops.learned_mask_mvdr(chunk, mask, cfg)
    torch.cuda.synchronize()
    t_mvdr = time.time() - tic
    print(f"MVDR Processing Time: {t_mvdr * 1000:.2f} ms")
    return out[0].cpu().numpy(), t_infer, t_mvdr


def main_deploy(input_path, model_path="mask_3.pth", arch="unet"):
    """full_audio.../inference.py:120-167: WAV in -> enhanced_<name>.wav in the cwd."""
    print(f"Processing {input_path}...")
    if not os.path.exists(model_path):
        print("Model not found.")
        return
    y_full, fs = wavio.read(input_path, dtype="float32")
    model = load_mask_model(model_path, arch)
    n = int(np.ceil(len(y_full) / (WIN_SIZE_SAMPLES // 2)))
    print(f"Audio Length: {len(y_full) / FS:.2f}s. Processing {n} sliding windows...")
    final_output = enhance_waveform(y_full, model, PRESETS["full_audio"], WIN_SIZE_SAMPLES)
    out_name = f"enhanced_{os.path.basename(input_path)}"
    wavio.write(out_name, final_output, fs)
    print(f"Saved: {out_name}")
    if "mixture_" in input_path and os.path.exists("target_ref_TEST.wav") and os.path.exists("interf_ref_TEST.wav"):
        tgt, _ = wavio.read("target_ref_TEST.wav")
        intf, _ = wavio.read("interf_ref_TEST.wav")
        n = min(len(final_output), len(tgt))
        sir, _ = calculate_metrics_manual(final_output[:n], tgt[:n], intf[:n])
        print(f"SIR Improvement: {sir:.2f} dB")
    return final_output

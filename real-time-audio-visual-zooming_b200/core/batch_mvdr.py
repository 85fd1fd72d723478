"""Drop-in for the operator surface of rt_av_zoom/core/tf_lite_version/inference.py:
`get_all_steering_vectors` (:53-81) and `batch_mvdr` (:85-179), plus `process_audio_file` (:245-391) with a torch
mask model in place of the TFLite interpreter (the .tflite blob is not shipped with the reference)."""
from __future__ import annotations

import dataclasses
import time

import numpy as np
import torch

from .. import ops, wavio
from ..config import MvdrConfig, PRESETS
from . import chunked

CONF = {"fs": 16000, "n_fft": 1024, "hop_len": 512, "d": 0.04, "c": 343.0, "train_seg_samples": 32000}  # :19-27 defaults
FS, N_FFT, HOP, D, C, WIN_SIZE = CONF["fs"], CONF["n_fft"], CONF["hop_len"], CONF["d"], CONF["c"], CONF["train_seg_samples"]
ANGLE_TARGET = 90.0
SIGMA = 1e-5
N_MICS = 2


def get_all_steering_vectors(f_bins, angle_deg, d, c):
    """(F,) bin frequencies -> (F, 2, 1) complex128 steering vectors (host float64, like the reference)."""
    f_bins = np.asarray(f_bins, dtype=np.float64)
    theta = np.deg2rad(angle_deg)
    tau1 = (d / 2) * np.cos(theta) / c
    tau2 = (d / 2) * np.cos(theta - np.pi) / c
    omega = 2 * np.pi * f_bins
    sv = np.stack([np.exp(-1j * omega * tau1), np.exp(-1j * omega * tau2)], axis=0)
    return np.expand_dims(sv.T, axis=-1)


def batch_mvdr(Y, mask, f_bins, d_vectors, sigma):
    """Y (2,F,T) complex, mask (F,T) target probability, d_vectors (F,2,1) -> beamformed (F,T) complex.
    Noise weight 1 - mask with 1e-10 inside the sqrt, normaliser + 1e-6, loading sigma, w eps 1e-10, no high-pass
    (tf_lite_version/inference.py:97-179).  numpy in -> numpy out, CUDA tensors in -> CUDA tensor out."""
    cfg = dataclasses.replace(PRESETS["tf_lite"], sigma=float(sigma))
    is_np = isinstance(Y, np.ndarray)
    Yt = torch.as_tensor(Y).cuda() if is_np else Y
    wide = Yt.dtype == torch.complex128            # complex128 in (scipy's STFT of float64 audio) -> float64 operators
    if not wide:
        Yt = Yt.to(torch.complex64)
    mt = torch.as_tensor(mask).to(Yt.device)
    noise_w = (1.0 - mt).to(torch.float64 if wide else torch.float32)
    dv = torch.as_tensor(np.asarray(d_vectors)).to(Yt.device, Yt.dtype).reshape(-1, 2)
    R = ops.masked_covariance(Yt, noise_w, sqrt_eps=cfg.sqrt_eps, norm_eps=cfg.norm_eps, packed=True)
    w = ops.mvdr_weights(R, dv, cfg)
    S = ops.beamform(w, Yt)
    return S.cpu().numpy() if is_np else S


class TFLiteBeamformer:
    """Name and call contract of tf_lite_version/inference.py:185-241: `predict_mask(log_mag (F,T), ipd (F,T)) -> (F,T)`.
    The interpreter is a torch module (`model=`, or a TorchScript file at `model_path`) that receives the reference's
    input tensor (1, F, T, 2) float32 NHWC = [log_mag, ipd]."""

    def __init__(self, model_path="mask_estimator.tflite", model=None):
        if model is None:
            model = torch.jit.load(model_path, map_location="cuda" if torch.cuda.is_available() else "cpu")
        self.model = model

    def predict_mask(self, log_mag, ipd):
        is_np = isinstance(log_mag, np.ndarray)
        lm = torch.as_tensor(log_mag, dtype=torch.float32)
        x = torch.stack([lm, torch.as_tensor(ipd, dtype=torch.float32).to(lm.device)], dim=-1)[None]
        with torch.no_grad():
            out = self.model(x).float().squeeze()
        return out.cpu().numpy() if is_np else out


def process_audio_file(input_path, output_path, model=None, model_path=None):
    """Chunked enhancement of a WAV file (:245-391): 2 s windows, 50 % overlap, count-averaged overlap-add, peak
    normalisation with 1e-9.  `model` is a torch mask estimator ((B,2,F,T) -> (B,F,T))."""
    y, sr = wavio.read(input_path, dtype="float32")
    if sr != FS:
        print("Warning: SR mismatch")
    print(f"Audio Duration:   {len(y) / FS:.2f}s")
    if model is None:
        model = chunked.load_mask_model(model_path)
    start = time.time()
    # per-window outputs stay un-normalised (the reference adds the raw iSTFT chunks, :351-373); only the final
    # waveform is divided by its peak (:375)
    chunk_cfg = dataclasses.replace(PRESETS["tf_lite"], peak_eps=None)
    enh = chunked.ChunkedEnhancer(chunk_cfg, WIN_SIZE, clip_to_input=True, final_peak_eps=1e-9)
    final = enh(chunked.to_planar(y), model)[0].cpu().numpy()
    proc_time = time.time() - start
    wavio.write(output_path, final, FS)
    print("-" * 40)
    print(f"Total Inference Time: {proc_time:.4f}s")
    print(f"Real-Time Factor:     {proc_time / (len(y) / FS):.4f}x")
    print(f"Saved to:             {output_path}")
    print("-" * 40)
    return final

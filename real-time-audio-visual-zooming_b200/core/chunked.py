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


def split_chunks(y_full: torch.Tensor, win: int):
    """(L, 2) -> chunks (n, 2, win): windows at stride win/2, n = ceil(L / stride), zero-padded tail
    (full_audio.../inference.py:137-147)."""
    L = y_full.shape[0]
    stride = win // 2
    n = int(np.ceil(L / stride))
    padded = torch.zeros((n * stride + win, y_full.shape[1]), dtype=y_full.dtype, device=y_full.device)
    padded[:L] = y_full
    idx = torch.arange(n, device=y_full.device)[:, None] * stride + torch.arange(win, device=y_full.device)[None, :]
    return padded[idx].permute(0, 2, 1).contiguous(), stride


def enhance_chunks(chunks: torch.Tensor, model, cfg: MvdrConfig) -> torch.Tensor:
    """process_chunk for a batch of windows: chunks (n, 2, win) -> (n, iSTFT length)."""
    X = ops.wave_features(chunks, cfg.n_fft, cfg.hop, "logmag_ipd")
    with torch.no_grad():
        mask = model(X).float().contiguous()
    return ops.learned_mask_mvdr(chunks, mask, cfg)


def overlap_add_chunks(outs: torch.Tensor, L: int, win: int, stride: int, buf_extra: int) -> torch.Tensor:
    """Count-averaged OLA (full_audio.../inference.py:135-156; Final_pipeline/src/inference.py:174-175,225-233).
    `buf_extra` = win for main_deploy (buffer of L + win, every chunk fits), 0 for the TFLite-era drivers (buffer of
    L: the last chunks are clipped)."""
    n, olen = outs.shape
    buf_len = L + buf_extra
    out_buf = torch.zeros(buf_len + win + olen, dtype=torch.float32, device=outs.device)
    cnt_buf = torch.zeros_like(out_buf)
    use = min(olen, win) if buf_extra else olen
    for i in range(n):                       # n is small (2 per second of audio); adds are tiny device ops
        s = i * stride
        m = min(use, buf_len - s)
        if m <= 0:
            continue
        out_buf[s:s + m] += outs[i, :m]
        cnt_buf[s:s + m] += 1.0
    cnt_buf = torch.where(cnt_buf == 0, torch.ones_like(cnt_buf), cnt_buf)
    return (out_buf / cnt_buf)[:L]


def enhance_waveform(y_full, model, cfg: MvdrConfig = PRESETS["full_audio"], win: int = WIN_SIZE_SAMPLES,
                     buf_extra: int | None = None) -> np.ndarray:
    """The body of main_deploy: (L, 2) waveform -> (L,) enhanced waveform."""
    if buf_extra is None:
        buf_extra = win
    y = torch.as_tensor(np.asarray(y_full, dtype=np.float32)).cuda()
    chunks, stride = split_chunks(y, win)
    outs = enhance_chunks(chunks, model, cfg)
    return overlap_add_chunks(outs, y.shape[0], win, stride, buf_extra).cpu().numpy()


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

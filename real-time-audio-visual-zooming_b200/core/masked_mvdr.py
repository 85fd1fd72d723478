"""Drop-in for rt_av_zoom/core/masked_mvdr.py: module constants, `get_steering_vector`,
`compute_hard_geometric_mask`, `main(output_dir_world)`."""
from __future__ import annotations

import os
import sys

import numpy as np

from .. import ops, wavio
from ..config import PRESETS

# --- 1. Constants (masked_mvdr.py:9-18) ---
FS = 16000
D = 0.01
C = 343.0
ANGLE_TARGET = 90.0
N_MICS = 2
SIGMA = 1e-7
N_FFT = 512
N_HOP = 256  # passed to scipy as `noverlap`; the hop happens to equal it (masked_mvdr.py:18,76)


def get_steering_vector(angle_deg, f, d, c):
    """masked_mvdr.py:22-35 -> (2, 1) complex128.  Two complex exponentials: host float64, like the reference
    (the batched per-bin form lives in ops.steering_vectors)."""
    theta_rad = np.deg2rad(angle_deg)
    tau_m1 = (d / 2) * np.cos(0.0) * np.cos(theta_rad - 0) / c
    tau_m2 = (d / 2) * np.cos(0.0) * np.cos(theta_rad - np.pi) / c
    omega = 2 * np.pi * f
    return np.array([[np.exp(-1j * omega * tau_m1)], [np.exp(-1j * omega * tau_m2)]], dtype=complex)


def compute_hard_geometric_mask(Y_stft, freqs):
    """masked_mvdr.py:37-46: (2,F,T) spectrum -> (F,T) mask, 1.0 where |angle(Y1) - angle(Y2)| > 0 else 0.01."""
    return ops.geometric_mask(Y_stft)


def enhance(y):
    """The arithmetic of main() (masked_mvdr.py:76-128) on an (L, 2) recording -> peak-normalised (n,) float64 waveform.
    sigma = 1e-7 on a near-rank-1 covariance amplifies STFT rounding by ~1e4, so this site runs on the float64
    operators (avz_*_f64), as the reference's complex128 arrays do."""
    import torch
    cfg = PRESETS["masked_mvdr"]
    yt = torch.from_numpy(np.ascontiguousarray(np.asarray(y, dtype=np.float64).T)).cuda()
    Y = ops.stft(yt, cfg.n_fft, cfg.hop)                                   # complex128
    mask_noise = ops.geometric_mask(Y)
    R = ops.masked_covariance(Y, mask_noise, cfg.sqrt_eps, cfg.norm_eps, packed=True)
    w = ops.mvdr_weights(R, ops.steering_vectors(cfg, Y.device, wide=True), cfg)
    s_out = ops.istft(ops.beamform(w, Y), cfg.n_fft, cfg.hop)
    return ops.peak_normalise(s_out[None], None, cfg.peak_eps)[0].cpu().numpy()


def main(output_dir_world):
    """masked_mvdr.py:50-135: <dir>/mixture_3_sources.wav -> <dir>/../MVDR_Outputs/output_masked_mvdr.wav."""
    if not output_dir_world or not os.path.exists(output_dir_world):
        print(f"ERROR: Invalid directory provided: {output_dir_world}")
        return
    print("--- 2. Masked MVDR Processing ---")
    input_file = os.path.join(output_dir_world, "mixture_3_sources.wav")
    if not os.path.exists(input_file):
        print(f"Error: {input_file} not found.")
        return
    run_root_dir = os.path.dirname(output_dir_world)
    mvdr_output_dir = os.path.join(run_root_dir, "MVDR_Outputs")
    os.makedirs(mvdr_output_dir, exist_ok=True)

    y, fs = wavio.read(input_file, dtype="float64")
    print("Calculating Hard Phase Mask...")
    print("Computing Weighted Noise Covariance...")
    print(f"Beamforming with SIGMA={SIGMA}...")
    s_out = enhance(y)
    wav_out_path = os.path.join(mvdr_output_dir, "output_masked_mvdr.wav")
    wavio.write(wav_out_path, s_out, fs)
    print("Done.")
    print(f"Saved outputs to: {mvdr_output_dir}")
    return wav_out_path


if __name__ == "__main__":
    if len(sys.argv) > 1:
        main(sys.argv[1])
    else:
        print("Usage: python masked_mvdr.py <simulation_output_directory>")

"""Drop-in for rt_av_zoom/core/oracle_reverb.py: `main(args)` with `--outdir --sigma --hp` (IRM post-filter)."""
from __future__ import annotations

import argparse
import dataclasses
import glob
import os

import numpy as np

from .. import ops, wavio
from ..config import PRESETS
from .masked_mvdr import get_steering_vector, D, C, N_MICS, FS, N_FFT, N_HOP  # noqa: F401

BASE_RESULTS_DIR = os.path.join(os.getcwd(), "simulation_results")
DEFAULT_OUTDIR = "simulation_results/ljspeech_reverb_20251130_215709"
ANGLE_TARGET = 90.0


def get_latest_run_dir():
    if not os.path.exists(BASE_RESULTS_DIR):
        return "simulation_results/latest_run"
    all_runs = sorted(glob.glob(os.path.join(BASE_RESULTS_DIR, "*_*_*")), reverse=True)
    return all_runs[0] if all_runs else "simulation_results/latest_run"


def enhance(y_mix, s_tgt, s_int, sigma, hp_cutoff):
    """The arithmetic of main() (oracle_reverb.py:76-160) on arrays: y_mix (2, L), s_tgt (L,), s_int (L,) float32 ->
    peak-normalised waveform (n,) float32.  IBM covariance, MVDR with sigma / hp, soft post-filter
    sqrt(Pt / (Pt + Pi + 1e-10)), x / (max|x| + 1e-9)."""
    import torch
    cfg = dataclasses.replace(PRESETS["oracle_debug"], sigma=float(sigma), hp_hz=float(hp_cutoff), post="mask",
                              peak_eps=1e-9)
    mix = torch.from_numpy(np.ascontiguousarray(y_mix, dtype=np.float32)).cuda()[None]
    tgt = torch.from_numpy(np.ascontiguousarray(s_tgt, dtype=np.float32)).cuda()[None]
    itf = torch.from_numpy(np.ascontiguousarray(s_int, dtype=np.float32)).cuda()[None]
    bits, Rp, _ = ops.ibm_covariance(mix, tgt, itf, cfg)
    w = ops.mvdr_weights(Rp, ops.steering_vectors(cfg, mix.device), cfg)
    mask_soft = ops.irm(ops.stft(tgt, cfg.n_fft, cfg.hop), ops.stft(itf, cfg.n_fft, cfg.hop))
    out, peak = ops.mvdr_apply(mix, w, cfg, mask=mask_soft)
    ops.peak_normalise(out, peak, cfg.peak_eps)
    return out[0].cpu().numpy()


def main(args):
    """oracle_reverb.py:41-178: IBM covariance on mixture_wpe.wav, MVDR with args.sigma / args.hp, soft post-filter
    sqrt(Pt / (Pt + Pi + 1e-10)), peak normalisation with 1e-9."""
    outdir, sigma, hp_cutoff = args.outdir, args.sigma, args.hp
    print("\n--- ORACLE OPTIMIZATION RUN ---")
    print(f"Directory:  {os.path.basename(outdir)}")
    print(f"Parameters: Sigma={sigma} | HP_Cutoff={hp_cutoff} Hz")
    if not os.path.exists(outdir):
        print(f"ERROR: Directory not found: {outdir}")
        return
    paths = [os.path.join(outdir, n) for n in ("mixture_wpe.wav", "target_reference.wav", "interference_reference.wav")]
    if not all(os.path.exists(p) for p in paths):
        print("CRITICAL ERROR: Audio files missing (mixture/target/interference).")
        return
    y_mix, _ = wavio.read(paths[0], dtype="float32")
    if y_mix.ndim > 1 and y_mix.shape[0] > y_mix.shape[1]:
        y_mix = y_mix.T
    s_tgt, _ = wavio.read(paths[1], dtype="float32")
    s_int, _ = wavio.read(paths[2], dtype="float32")
    print("Oracle Mask generated (Includes Reverb tails in Interference).")
    out = enhance(y_mix, s_tgt, s_int, sigma, hp_cutoff)
    out_path = os.path.join(outdir, "output_oracle_reverb.wav")
    wavio.write(out_path, out, FS)
    print(f"Saved: {out_path}")
    print("-----------------------------------")
    return out_path


if __name__ == "__main__":
    parser = argparse.ArgumentParser(description="Run Optimized Oracle MVDR")
    parser.add_argument("--outdir", type=str, default=DEFAULT_OUTDIR)
    parser.add_argument("--sigma", type=float, default=1e-3)
    parser.add_argument("--hp", type=float, default=100.0)
    main(parser.parse_args())

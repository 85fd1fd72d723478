"""Drop-in for rt_av_zoom/core/oracle_debug.py: `main()` (oracle IBM mask-MVDR on three WAV files)."""
from __future__ import annotations

import os

import numpy as np

from .. import ops, wavio
from ..config import PRESETS
from .masked_mvdr import get_steering_vector, D, C, N_MICS, FS, N_FFT, N_HOP  # noqa: F401  (oracle_debug.py:11-19)

ANGLE_TARGET = 90.0
SIGMA = 1
OUTDIR = "simulation_results/ljspeech_anechoic_20251130_154029"


def main():
    """oracle_debug.py:27-97: reads {OUTDIR}/mixture.wav, target_reference.wav, interference_reference.wav (relative to
    the cwd), writes {OUTDIR}/output_oracle.wav.  The whole of lines 42-94 runs as the fused two-pass GPU path."""
    print("--- ORACLE TEST: Can the code theoretically work? ---")
    if not os.path.exists(f"{OUTDIR}//target_reference.wav") or not os.path.exists(f"{OUTDIR}//interference_reference.wav"):
        print("Error: Reference files missing. Run world.py first.")
        return
    y_mix, _ = wavio.read(f"{OUTDIR}//mixture.wav", dtype="float32")
    s_tgt_ref, _ = wavio.read(f"{OUTDIR}//target_reference.wav", dtype="float32")
    s_int_ref, _ = wavio.read(f"{OUTDIR}//interference_reference.wav", dtype="float32")
    print("Oracle Mask created directly from ground truth files.")
    print("Computing Covariance...")
    print("Running MVDR...")
    print("Applying Aggressive Post-Filter...")
    s_out = ops.oracle_mask_mvdr(np.ascontiguousarray(y_mix.T), s_tgt_ref, s_int_ref, PRESETS["oracle_debug"])
    wavio.write(f"{OUTDIR}/output_oracle.wav", s_out, FS)
    print(f"Saved '{OUTDIR}/output_oracle.wav'.")
    return s_out


if __name__ == "__main__":
    main()

"""Mirror of Final_pipeline/src/inference.py: `get_steering_vector_single` (:16-26) and
`enhance_audio(run_name, input_path, model_path)` (:144-238).

The chunk loop (2 s windows, 50 % overlap, count-averaged OLA clipped to the input length, peak normalisation with
1e-9) is the reference's; the per-chunk beamformer here is the mask-driven MVDR of this project's hot path with the
Final_pipeline constants (n_fft 1024 / hop 512, d = 0.08, mic-0 pass-through below 200 Hz, post-filter S * mask).
The reference's `hybrid_hard_null_bf` (:28-98, an eigenvector null-steering beamformer, not MVDR) is the next row of
the scope table and is not built yet; calling it raises NotImplementedError."""
from __future__ import annotations

import os
import time

import numpy as np

from .. import wavio
from ..config import MvdrConfig
from ..core import chunked
from . import config

ANGLE_TARGET = 90.0
FREQ_BINS = (config.N_FFT // 2) + 1
N_MICS = 2

FINAL_CFG = MvdrConfig(fs=config.FS, n_fft=config.N_FFT, hop=config.HOP_LEN, mic_dist=config.MIC_DIST, c=config.C_SPEED,
                       sigma=1e-5, hp_hz=200.0, hp_mode="mic0", post="mask", peak_eps=None)


def get_steering_vector_single(f, angle_deg, d, c):
    """Steering vector of one bin, phase-normalised to mic 0 (Final_pipeline/src/inference.py:16-26) -> (2, 1)."""
    theta = np.deg2rad(angle_deg)
    tau1 = (d / 2) * np.cos(theta) / c
    tau2 = (d / 2) * np.cos(theta - np.pi) / c
    omega = 2 * np.pi * f
    v = np.array([[np.exp(-1j * omega * tau1)], [np.exp(-1j * omega * tau2)]])
    return v / (v[0] + 1e-10)


def hybrid_hard_null_bf(Y, mask, f_bins):
    raise NotImplementedError("hybrid hard-null beamformer (Final_pipeline/src/inference.py:28-98): next scope row, "
                              "see DESIGN.md; enhance_audio uses mask-driven MVDR")


def enhance_audio(run_name, input_path, model_path, model=None):
    """Reads `input_path`, writes {RESULTS_DIR}/{run_name}_results/{run_name}_enhanced.wav."""
    result_dir = os.path.join(config.RESULTS_DIR, f"{run_name}_results")
    os.makedirs(result_dir, exist_ok=True)
    output_path = os.path.join(result_dir, f"{run_name}_enhanced.wav")
    print(f"[INF] Processing {input_path}")
    print(f"[INF] Saving to  {output_path}")
    y, sr = wavio.read(input_path, dtype="float32")
    if sr != config.FS:
        print(f"Warning: SR mismatch. Input: {sr}, Config: {config.FS}")
    if y.ndim == 1:
        print("Error: Input is mono. Requires 2 channels.")
        return
    if model is None:
        try:
            model = chunked.load_mask_model(model_path if model_path and os.path.exists(model_path) else None)
        except Exception as e:  # same contract as the reference: report and give up on this run
            print(f"Failed to load mask model: {e}")
            return
    start_time = time.time()
    final = chunked.enhance_waveform(y, model, FINAL_CFG, win=config.WIN_SIZE, buf_extra=0)
    print(f"Total processing time: {time.time() - start_time:.3f}s")
    final = final / (np.max(np.abs(final)) + 1e-9)
    wavio.write(output_path, final, config.FS)
    return output_path

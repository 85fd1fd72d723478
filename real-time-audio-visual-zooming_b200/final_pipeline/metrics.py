"""Mirror of Final_pipeline/src/metrics.py: `calculate_osnr_osir` (:102-123) and `evaluate_run` (:125-206) limited to
the projection scores (STOI / PESQ need pystoi / pesq, absent and out of scope)."""
from __future__ import annotations

import os

import numpy as np

from .. import wavio
from ..core.metrics import calculate_osnr_osir  # noqa: F401
from . import config


def load_and_align(run_name):
    sim = os.path.join(config.SIM_DIR, run_name)
    res = os.path.join(config.RESULTS_DIR, f"{run_name}_results", f"{run_name}_enhanced.wav")
    try:
        s_est, _ = wavio.read(res)
        s_tgt, _ = wavio.read(os.path.join(sim, "target_reference.wav"))
        s_int, _ = wavio.read(os.path.join(sim, "interference_reference.wav"))
        s_mix, _ = wavio.read(os.path.join(sim, "mixture.wav"))
    except FileNotFoundError as e:
        print(f"[EVAL] Error: Missing file - {e}")
        return None, None, None, None
    if s_mix.ndim > 1:
        s_mix = s_mix[:, 0]
    n = min(len(s_est), len(s_tgt), len(s_int), len(s_mix))
    return (s_est[:n].astype(np.float64), s_tgt[:n].astype(np.float64), s_int[:n].astype(np.float64),
            s_mix[:n].astype(np.float64))


def evaluate_run(run_name):
    s_est, s_tgt, s_int, s_mix = load_and_align(run_name)
    if s_est is None:
        return None
    osinr_in, osir_in = calculate_osnr_osir(s_mix, s_tgt, s_int)
    osinr_out, osir_out = calculate_osnr_osir(s_est, s_tgt, s_int)
    report = {"run": run_name, "SIR_in": osir_in, "SIR_out": osir_out, "SIR_Imp": osir_out - osir_in,
              "SINR_in": osinr_in, "SINR_out": osinr_out}
    print(f"[EVAL] {run_name}: SIR {osir_in:.2f} -> {osir_out:.2f} dB (+{osir_out - osir_in:.2f}), "
          f"SINR {osinr_in:.2f} -> {osinr_out:.2f} dB")
    return report

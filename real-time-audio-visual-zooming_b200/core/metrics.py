"""Drop-in for the projection scores: `calculate_osnr_osir` (Final_pipeline/src/metrics.py:102-123) and
`calculate_metrics_manual` (scripts/run_metrics.py:6-36).  Both run as one GPU reduction kernel (float64)."""
from __future__ import annotations

import numpy as np

from .. import ops


def _scores(output, target, interferer):
    return ops.sir_scores(np.asarray(output, np.float32), np.asarray(target, np.float32),
                          np.asarray(interferer, np.float32))


def calculate_osnr_osir(output, target, interferer):
    """-> (OSINR, OSIR) in dB."""
    sc = _scores(output, target, interferer)
    return float(sc[0]), float(sc[1])


def calculate_metrics_manual(output_signal, target_ref, interf_ref):
    """-> (sdr, sir) in dB (the output is scaled to unit norm first, as run_metrics.py does)."""
    sc = _scores(output_signal, target_ref, interf_ref)
    return float(sc[2]), float(sc[3])

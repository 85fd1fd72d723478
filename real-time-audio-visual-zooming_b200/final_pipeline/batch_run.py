"""Mirror of Final_pipeline/batch_run.py: `run_batch(n_runs, start_idx=0, n_interferers=2)` - simulate, enhance,
evaluate per run, exceptions swallowed per run (batch_run.py:12-49)."""
from __future__ import annotations

import argparse
import os

from . import config, inference, metrics, simulation


def run_batch(n_runs, start_idx=0, n_interferers=2, model=None):
    model_path = os.path.join("models", "mask_estimator_phy.pth")
    print(f"=== STARTING BATCH RUN: {n_runs} Iterations ===")
    reports = []
    for i in range(start_idx, start_idx + n_runs):
        run_name = f"batch_test_{i:03d}"
        try:
            mix_path = simulation.generate_scene(run_name=run_name, dataset="synthetic", reverb=False,
                                                 n_interferers=n_interferers, snr_target=50)
            if not mix_path:
                continue
            inference.enhance_audio(run_name=run_name, input_path=mix_path, model_path=model_path, model=model)
            reports.append(metrics.evaluate_run(run_name))
        except Exception as e:
            print(f"\n[ERROR] Failed on {run_name}: {e}")
            continue
    return reports


if __name__ == "__main__":
    parser = argparse.ArgumentParser()
    parser.add_argument("--n", type=int, default=10, help="Number of runs")
    parser.add_argument("--start", type=int, default=0, help="Start index for naming")
    parser.add_argument("--interferers", type=int, default=2, help="Number of interferers")
    args = parser.parse_args()
    run_batch(args.n, args.start, args.interferers)

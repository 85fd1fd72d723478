"""Mirror of Final_pipeline/run.py: `python -m ... run {sim,inf,eval,full} --name NAME` (run.py:5-59)."""
from __future__ import annotations

import argparse
import os

from . import config, inference, metrics, simulation


def main(argv=None):
    parser = argparse.ArgumentParser(description="Neural Beamforming Pipeline")
    parser.add_argument("mode", choices=["sim", "inf", "eval", "full"], help="Action: sim, inf, eval, or full")
    parser.add_argument("--name", type=str, required=True, help="Run Name (e.g., 'test1'). Used to find folders.")
    parser.add_argument("--reverb", action="store_true", default=True)
    parser.add_argument("--no-reverb", action="store_false", dest="reverb")
    parser.add_argument("--dataset", default="ljspeech")
    parser.add_argument("--n", type=int, default=1)
    parser.add_argument("--snr", type=int, default=5)
    args = parser.parse_args(argv)
    sim_folder = os.path.join(config.SIM_DIR, args.name)
    mixture_path = os.path.join(sim_folder, "mixture.wav")
    model_path = os.path.join(config.PROJECT_ROOT, "models", "mask_estimator_phy.pth")
    if args.mode in ["sim", "full"]:
        print(f"\n--- 1. BUILDING WORLD: {args.name} ---")
        if os.path.exists(sim_folder) and args.mode == "sim":
            print(f"Warning: {sim_folder} exists.")
        simulation.generate_scene(run_name=args.name, dataset=args.dataset, reverb=args.reverb, n_interferers=args.n,
                                  snr_target=args.snr)
    if args.mode in ["inf", "full"]:
        print(f"\n--- 2. RUNNING INFERENCE: {args.name} ---")
        if not os.path.exists(mixture_path):
            print("Error: Mixture not found. Run 'sim' first.")
            return
        inference.enhance_audio(run_name=args.name, input_path=mixture_path, model_path=model_path)
    if args.mode in ["eval", "inf", "full"]:
        print(f"\n--- 3. EVALUATING RESULTS: {args.name} ---")
        metrics.evaluate_run(args.name)


if __name__ == "__main__":
    main()

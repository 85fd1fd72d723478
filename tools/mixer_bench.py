"""Time the on-device far-field mixer at BASELINE config 2's shape (1024 x 4 sources x 4 s).  python tools/mixer_bench.py [B] [S] [L]"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import avzoom  # noqa: E402
from avzoom import ops  # noqa: E402


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
    S = int(sys.argv[2]) if len(sys.argv) > 2 else 4
    L = int(sys.argv[3]) if len(sys.argv) > 3 else 64000
    src = torch.randn((B, S, L), device="cuda")
    angles = [90.0, 40.0, 130.0, 65.0, 155.0, 20.0, 110.0, 75.0][:S]
    th = np.deg2rad(angles)
    delays = np.stack([0.02 * np.cos(th) / 343.0, 0.02 * np.cos(th - np.pi) / 343.0], axis=1)
    avzoom._lib.load().avz_profile_enable(0)
    for _ in range(2):
        ops.far_field_mix(src, delays)
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    reps = 5
    ev[0].record()
    for _ in range(reps):
        ops.far_field_mix(src, delays)
    ev[1].record()
    torch.cuda.synchronize()
    ms = ev[0].elapsed_time(ev[1]) / reps
    print(json.dumps({"B": B, "S": S, "L": L, "ms_per_batch": round(ms, 3),
                      "audio_s_per_s": round(B * L / 16000.0 / (ms * 1e-3), 1),
                      "note": "includes the torch.empty allocations of outputs and workspace"}))


if __name__ == "__main__":
    main()

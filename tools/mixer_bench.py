"""Time the on-device far-field mixer at BASELINE config 2's shape (1024 x 4 sources x 4 s): the cluster-resident kernel
(`avz_farfield_mix_f32`) and the multi-pass transform through HBM (`avz_farfield_mix_passes_f32`), outputs and workspace
allocated once.  python tools/mixer_bench.py [B] [S] [L]

With an experiment build (`tools/build_exp.sh`, AVZ_LIB=...) and AVZ_MIX_PHASES=1 the cluster kernel is also timed
stopped after its load / forward stages / exchange step (AVZ_MIX_DBG = 1, 2, 3): the differences are the phase costs."""
import ctypes as C
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import avzoom  # noqa: E402
from avzoom import _lib  # noqa: E402


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
    S = int(sys.argv[2]) if len(sys.argv) > 2 else 4
    L = int(sys.argv[3]) if len(sys.argv) > 3 else 64000
    lib = _lib.load()
    src = torch.randn((B, S, L), device="cuda")
    angles = [90.0, 40.0, 130.0, 65.0, 155.0, 20.0, 110.0, 75.0][:S]
    th = np.deg2rad(angles)
    delays = np.ascontiguousarray(np.stack([0.02 * np.cos(th) / 343.0, 0.02 * np.cos(th - np.pi) / 343.0], axis=1))
    ws = torch.empty(int(lib.avz_farfield_mix_ws_bytes(B, S, L)), dtype=torch.uint8, device="cuda")
    mix = torch.empty((B, 2, L), device="cuda")
    tgt = torch.empty((B, L), device="cuda")
    itf = torch.empty((B, L), device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    dp = delays.ctypes.data_as(C.POINTER(C.c_double))

    def run(fn):
        _lib.check(fn(src.data_ptr(), dp, B, S, L, 16000.0, 1e-9, mix.data_ptr(), tgt.data_ptr(), itf.data_ptr(),
                      ws.data_ptr(), st), "mix")

    def time_ms(fn, reps=10):
        for _ in range(3):
            run(fn)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            run(fn)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    out = {"B": B, "S": S, "L": L}
    ms = time_ms(lib.avz_farfield_mix_f32)
    out["cluster_or_default_ms"] = round(ms, 3)
    out["audio_s_per_s"] = round(B * L / 16000.0 / (ms * 1e-3), 1)
    out["multi_pass_ms"] = round(time_ms(lib.avz_farfield_mix_passes_f32), 3)
    out["io_bytes"] = int(src.numel() * 4 + mix.numel() * 4 + tgt.numel() * 4 + itf.numel() * 4)
    out["GBps_of_io"] = round(out["io_bytes"] / (ms * 1e-3) / 1e9, 1)
    if os.environ.get("AVZ_MIX_PHASES"):
        ph = {}
        for d, name in ((1, "load+store"), (2, "+forward stages"), (3, "+exchange step")):
            os.environ["AVZ_MIX_DBG"] = str(d)
            ph[name] = round(time_ms(lib.avz_farfield_mix_f32), 3)
        os.environ["AVZ_MIX_DBG"] = "0"
        ph["+inverse stages, peak, normalised store"] = round(time_ms(lib.avz_farfield_mix_f32), 3)
        out["cluster_phases_cumulative_ms"] = ph
    print(json.dumps(out))


if __name__ == "__main__":
    main()

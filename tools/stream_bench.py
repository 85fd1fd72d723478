"""BASELINE config 4: per-hop latency of the streaming step across S concurrent 2-mic streams (default 4096 streams,
1000 hops).  Device time per hop from CUDA events; also through a CUDA graph of one step (launch overhead removed).
usage: python tools/stream_bench.py [S] [hops]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import avzoom  # noqa: E402
from avzoom import stream  # noqa: E402

S = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
hops = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
eng = stream.MvdrStream(S, avzoom.PRESETS["baseline_oracle"], lam=0.95)
x = torch.randn((S, 2, 128), device="cuda") * 0.1
m = torch.rand((S, 257), device="cuda")
for _ in range(20):
    eng.step(x, m)
torch.cuda.synchronize()
ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(hops)]
for a, b in ev:
    a.record()
    eng.step(x, m)
    b.record()
torch.cuda.synchronize()
t = np.array([a.elapsed_time(b) for a, b in ev]) * 1e3   # microseconds
res = {"config": "BASELINE config 4: streaming recursive-covariance MVDR, n_fft 512 hop 128, lambda 0.95",
       "streams": S, "hops": hops, "hop_us_p50": float(np.percentile(t, 50)), "hop_us_p99": float(np.percentile(t, 99)),
       "hop_us_mean": float(t.mean()), "stream_hops_per_s": S / (t.mean() * 1e-6),
       "real_time_factor": 8000.0 / float(np.percentile(t, 99)),   # a hop is 8 ms of audio
       "state_bytes_per_stream": int(eng.lib.avz_stream_state_bytes(1)),
       "algorithmic_bytes_per_stream_hop": 2 * int(eng.lib.avz_stream_state_bytes(1)) + 1024 + 512 + 257 * 4}
res["achieved_GBps"] = res["algorithmic_bytes_per_stream_hop"] * S / (t.mean() * 1e-6) / 1e9
print(json.dumps(res))

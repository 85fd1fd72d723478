"""BASELINE config 1 (the reference's own CPU-runnable case): ONE 5 s 2-mic mixture, 1 target + 2 interferers,
n_fft 512 / hop 128, oracle IBM mask-MVDR.  A single utterance is launch-bound (7 kernels + 2 memsets of a few
microseconds each), so the step is also replayed as a CUDA graph.  Device time from CUDA events (the CPU figure to hold against it is bench.py's cpu_baseline: ~6.7 ms per audio-second
and core).
usage: python tools/c1_latency.py [B]"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import avzoom  # noqa: E402
from avzoom import pipeline, synth  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
cfg = avzoom.PRESETS["baseline_oracle"]
dev = torch.device("cuda", 0)
mix, tgt, itf = synth.make_batch(1, B, 5.0, 2)
L = mix.shape[-1]
mix_d, tgt_d, itf_d = (torch.from_numpy(a).to(dev) for a in (mix, tgt, itf))
eng = pipeline.OracleMvdr(cfg, B, L, dev)
for _ in range(5):
    eng.run(mix_d, tgt_d, itf_d)
torch.cuda.synchronize()
ref_out = eng.out.clone()


def timed(fn, iters=200):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters * 1e3, (time.perf_counter() - t0) / iters * 1e6


eager_dev_us, eager_wall_us = timed(lambda: eng.run(mix_d, tgt_d, itf_d))

side = torch.cuda.Stream()
side.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(side):
    eng.run(mix_d, tgt_d, itf_d)
torch.cuda.current_stream().wait_stream(side)
graph = torch.cuda.CUDAGraph()
with torch.cuda.graph(graph):
    eng.run(mix_d, tgt_d, itf_d)
graph.replay()
torch.cuda.synchronize()
same = bool(torch.equal(eng.out, ref_out))
graph_dev_us, graph_wall_us = timed(graph.replay)

res = {"config": "BASELINE config 1: one 5 s mixture (1 target + 2 interferers), oracle IBM mask-MVDR, n_fft 512 hop 128",
       "utterances": B, "samples": L, "eager_us": eager_dev_us, "eager_host_wall_us": eager_wall_us,
       "cuda_graph_us": graph_dev_us, "cuda_graph_host_wall_us": graph_wall_us, "graph_output_bit_identical": same,
       "real_time_factor_graph": B * 5.0 / (graph_dev_us * 1e-6)}
res["kernel_us_eager"] = {k: round(v * 1e3, 2) for k, v in eng.time_each_kernel(mix_d, tgt_d, itf_d, 20).items()}
print(json.dumps(res))

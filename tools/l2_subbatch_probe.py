"""Does the kept spectrum stay in L2 if a step is cut into sub-batches?  

A step over 1024 utterances runs as 1024/sb sub-batches, each a full pass of the seven kernels over `sb` utterances with
its own small workspace (kept spectrum sb x 2.05 MB, reused by every later sub-batch of the same engine, so its lines can
live in the 126 MB L2 between pass A and pass B), round-robin over `depth` engines on `depth` streams; the whole step is
one CUDA graph (no per-launch host cost).  Prints ms per step for each (sb, depth) and checks bit-identity with the
one-launch-per-kernel step.
usage: python tools/l2_subbatch_probe.py [B] [--ncu SB DEPTH]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import avzoom  # noqa: E402
from avzoom import pipeline, synth  # noqa: E402

argv = [a for a in sys.argv[1:]]
ncu = None
if "--ncu" in argv:
    i = argv.index("--ncu")
    ncu = (int(argv[i + 1]), int(argv[i + 2]))
    del argv[i:i + 3]
B = int(argv[0]) if argv else 1024
cfg = avzoom.PRESETS["baseline_oracle"]
dev = torch.device("cuda", 0)
distinct = min(B, 128)
mix, tgt, itf = synth.make_batch(2, distinct, 4.0, 3, workers=min(16, os.cpu_count() or 1))
rep = B // distinct
mix, tgt, itf = np.tile(mix, (rep, 1, 1)), np.tile(tgt, (rep, 1)), np.tile(itf, (rep, 1))
L = mix.shape[-1]
mix_d, tgt_d, itf_d = (torch.from_numpy(a).to(dev) for a in (mix, tgt, itf))

full = pipeline.OracleMvdr(cfg, B, L, dev)
for _ in range(3):
    full.run(mix_d, tgt_d, itf_d)
torch.cuda.synchronize()
ref = full.out.clone()
out = torch.empty_like(ref)


def timed(fn, iters=20):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


def build(sb, depth):
    engines = [pipeline.OracleMvdr(cfg, sb, L, dev) for _ in range(depth)]
    streams = [torch.cuda.Stream(dev) for _ in range(depth)]
    nsub = B // sb

    def step():
        cur = torch.cuda.current_stream()
        for s in streams:
            s.wait_stream(cur)
        for i in range(nsub):
            e, s = engines[i % depth], streams[i % depth]
            lo, hi = i * sb, (i + 1) * sb
            with torch.cuda.stream(s):
                e.run(mix_d[lo:hi], tgt_d[lo:hi], itf_d[lo:hi], out=out[lo:hi])
        for s in streams:
            cur.wait_stream(s)
    return step, engines


res = {"B": B, "full_ms": timed(lambda: full.run(mix_d, tgt_d, itf_d)), "runs": []}
print(json.dumps({"full_ms": res["full_ms"]}), flush=True)
if ncu is not None:
    step, keep = build(*ncu)
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    sys.exit(0)
for sb in (16, 24, 32, 48, 64, 128, 256):
    if B % sb:
        continue
    for depth in (1, 2, 3):
        step, keep = build(sb, depth)
        step()
        torch.cuda.synchronize()
        same = bool(torch.equal(out, ref))
        side = torch.cuda.Stream(dev)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            step()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            step()
        ms = timed(g.replay)
        r = {"sub_batch": sb, "depth": depth, "kept_spectrum_MB_in_flight": depth * sb * 2.05, "graph_ms": ms,
             "bit_identical": same}
        res["runs"].append(r)
        print(json.dumps(r), flush=True)
        del g, step, keep
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/l2_subbatch_probe.json", "w"), indent=1)

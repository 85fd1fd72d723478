"""Does running sub-batches of the step on concurrent streams raise throughput?  (issue-bound k512_ibm next to the
HBM-heavy k512_cov / k512_apply of another sub-batch.)   python tools/overlap_probe.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import avzoom  # noqa: E402
from avzoom import pipeline, synth  # noqa: E402

B, L = 1024, 64000
cfg = avzoom.PRESETS["baseline_oracle"]
mix, tgt, itf = synth.make_batch(2, 16, 4.0, 3)
mix = torch.from_numpy(mix).cuda().repeat(B // 16, 1, 1).contiguous()
tgt = torch.from_numpy(tgt).cuda().repeat(B // 16, 1).contiguous()
itf = torch.from_numpy(itf).cuda().repeat(B // 16, 1).contiguous()
dev = mix.device


def bench(nsub, nstreams, stagger, steps=10):
    sb = B // nsub
    engs = [pipeline.OracleMvdr(cfg, sb, L, dev) for _ in range(nsub)]
    streams = [torch.cuda.Stream(dev) for _ in range(nstreams)]
    cur = torch.cuda.current_stream()

    def step():
        for s in streams:
            s.wait_stream(cur)
        if stagger:   # stage-major issue order: pass A of every sub-batch first, then the rest
            for i, e in enumerate(engs):
                with torch.cuda.stream(streams[i % nstreams]):
                    lo = i * sb
                    e.pass_a(mix[lo:lo + sb], tgt[lo:lo + sb], itf[lo:lo + sb])
            for i, e in enumerate(engs):
                with torch.cuda.stream(streams[i % nstreams]):
                    e.weights(); e.pass_b(mix[i * sb:(i + 1) * sb]); e.normalise()
        else:
            for i, e in enumerate(engs):
                with torch.cuda.stream(streams[i % nstreams]):
                    lo = i * sb
                    e.run(mix[lo:lo + sb], tgt[lo:lo + sb], itf[lo:lo + sb])
        for s in streams:
            cur.wait_stream(s)

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    print(f"nsub={nsub} streams={nstreams} stagger={stagger}: {ms:.3f} ms/step  {B * 4.0 / ms / 1e3:.3f} M audio-s/s", flush=True)
    del engs


for nsub, ns, st in ((1, 1, False), (2, 2, False), (4, 4, False), (4, 2, False), (8, 4, False), (8, 8, False), (4, 4, True), (8, 4, True), (16, 4, False)):
    bench(nsub, ns, st)

"""Do concurrent streams raise throughput?  (issue-bound k512_ibm next to the latency-bound k512_cov / k512_apply.)
Mode 1: sub-batches of one step on concurrent streams.  Mode 2: whole batches of consecutive steps on two streams
(steady-state pipelining of a serving loop).   python tools/overlap_probe.py"""
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


def timed(step, steps):
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def sub_batches(nsub, nstreams):
    sb = B // nsub
    engs = [pipeline.OracleMvdr(cfg, sb, L, dev) for _ in range(nsub)]
    streams = [torch.cuda.Stream(dev) for _ in range(nstreams)]
    cur = torch.cuda.current_stream()

    def step():
        for s in streams:
            s.wait_stream(cur)
        for i, e in enumerate(engs):
            with torch.cuda.stream(streams[i % nstreams]):
                lo = i * sb
                e.run(mix[lo:lo + sb], tgt[lo:lo + sb], itf[lo:lo + sb])
        for s in streams:
            cur.wait_stream(s)

    ms = timed(step, 10)
    print(f"sub-batches: nsub={nsub} streams={nstreams}: {ms:.3f} ms/step  {B * 4.0 / ms / 1e3:.3f} M audio-s/s", flush=True)


def whole_batches(nstreams, stagger):
    """Consecutive steps alternate between `nstreams` engines/streams; `stagger` delays every stream but the first by
    one pass A so that unlike kernels meet on the SMs."""
    engs = [pipeline.OracleMvdr(cfg, B, L, dev) for _ in range(nstreams)]
    streams = [torch.cuda.Stream(dev) for _ in range(nstreams)]
    cur = torch.cuda.current_stream()
    rounds = 10

    def step():   # = nstreams steps of work
        for s in streams:
            s.wait_stream(cur)
        if stagger:
            for i, e in enumerate(engs):
                with torch.cuda.stream(streams[i]):
                    e.pass_a(mix, tgt, itf)
            for i, e in enumerate(engs):
                with torch.cuda.stream(streams[i]):
                    e.weights(); e.pass_b(mix); e.normalise()
        else:
            for i, e in enumerate(engs):
                with torch.cuda.stream(streams[i]):
                    e.run(mix, tgt, itf)
        for s in streams:
            cur.wait_stream(s)

    ms = timed(step, rounds) / nstreams
    print(f"whole batches: streams={nstreams} stagger={stagger}: {ms:.3f} ms/step  {B * 4.0 / ms / 1e3:.3f} M audio-s/s", flush=True)


sub_batches(1, 1)
sub_batches(2, 2)
sub_batches(4, 4)
whole_batches(1, False)
whole_batches(2, False)
whole_batches(2, True)
whole_batches(3, False)

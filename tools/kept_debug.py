import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, avzoom
from avzoom import synth, pipeline
cfg = avzoom.PRESETS["baseline_oracle"]
B, dur = 300, 0.25
mix, tgt, itf = synth.make_batch(4, B, dur, 2)
mix_d, tgt_d, itf_d = (torch.from_numpy(a).cuda() for a in (mix, tgt, itf))
ek = pipeline.OracleMvdr(cfg, B, mix.shape[-1], mix_d.device, keep_spectrum=True)
er = pipeline.OracleMvdr(cfg, B, mix.shape[-1], mix_d.device, keep_spectrum=False)
for trial in range(3):
    a = ek.run(mix_d, tgt_d, itf_d).clone(); b = er.run(mix_d, tgt_d, itf_d).clone()
    d = (a != b)
    print("trial", trial, "bits equal", torch.equal(ek.bits, er.bits), "R equal", torch.equal(ek.R, er.R), "w equal", torch.equal(ek.w, er.w),
          "peak equal", torch.equal(ek.peak, er.peak), "n diff samples", int(d.sum()), "utts", torch.nonzero(d.any(1)).flatten().tolist()[:10])
    if d.any():
        u = int(torch.nonzero(d.any(1))[0])
        idx = torch.nonzero(d[u]).flatten()
        print("  utt", u, "first/last diff sample", int(idx[0]), int(idx[-1]), "count", len(idx), "max abs diff", float((a[u]-b[u]).abs().max()), "peak k/r", float(ek.peak[u]), float(er.peak[u]))

"""What bounds k512_cov / k512_apply?  Kernel times (CUDA events) of the same step with the kept-spectrum stores on and
off (OracleMvdr(keep_spectrum=...)); run it again with AVZ_LIB pointing at a build made with -DAVZ_COV_NOFFT to take the
transform out of k512_cov instead (profiles/README.md, "DRAM-bound or not").
usage: [AVZ_LIB=...] python tools/cov_bound_probe.py [B]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import avzoom  # noqa: E402
from avzoom import pipeline, synth  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
mix, tgt, itf = synth.make_batch(2, 128, 4.0, 3, workers=min(16, os.cpu_count() or 1))
rep = B // 128
mix, tgt, itf = (torch.from_numpy(np.tile(a, (rep,) + (1,) * (a.ndim - 1))).cuda() for a in (mix, tgt, itf))
res = {"lib": os.environ.get("AVZ_LIB", "release"), "B": B}
for keep in (True, False):
    e = pipeline.OracleMvdr(avzoom.PRESETS["baseline_oracle"], B, mix.shape[-1], mix.device, keep_spectrum=keep)
    k = e.time_each_kernel(mix, tgt, itf, iters=10)
    res["kept_spectrum" if keep else "recompute"] = {a: round(b, 4) for a, b in k.items()}
    del e
print(json.dumps(res))

"""Per-kernel times of the oracle engine's opt-in variants on BASELINE config 2 (1024 x 4 s):
default (store skipping), sparse kept spectrum, fused persistent kernel, folded weights.  python tools/engine_variants.py [B]"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import avzoom  # noqa: E402
from avzoom import pipeline, synth  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
cfg = avzoom.PRESETS["baseline_oracle"]
mix, tgt, itf = synth.make_batch(2, B, 4.0, 3, workers=min(16, os.cpu_count() or 1))
mix, tgt, itf = (torch.from_numpy(a).cuda() for a in (mix, tgt, itf))
lib = avzoom._lib.load()
res = {}
for name, kw in (("default (store skipping)", {}), ("dense stores", {"skip_masked_stores": False}),
                 ("sparse kept spectrum", {"sparse_spectrum": True}), ("fused persistent kernel", {"fused": True}),
                 ("weights folded into pass A", {"fold_weights": True})):
    eng = pipeline.OracleMvdr(cfg, B, mix.shape[-1], mix.device, **kw)
    for _ in range(3):
        eng.run(mix, tgt, itf)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        eng.run(mix, tgt, itf)
    e1.record()
    torch.cuda.synchronize()
    res[name] = round(e0.elapsed_time(e1) / 20, 4)
    del eng
    torch.cuda.empty_cache()
print(json.dumps({"B": B, "ms_per_step_single_stream": res}))

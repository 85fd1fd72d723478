"""A/B of the cluster-fused peak normalisation (OracleMvdr(fused_norm=True)) against the separate k_peak_normalise pass.
python tools/fused_norm_ab.py"""
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
outs = {}
for fused in (False, True, False, True):
    eng = pipeline.OracleMvdr(cfg, B, L, mix.device, fused_norm=fused)
    for _ in range(3):
        eng.run(mix, tgt, itf)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        eng.run(mix, tgt, itf)
    e1.record()
    torch.cuda.synchronize()
    k = eng.time_each_kernel(mix, tgt, itf)
    print(f"fused_norm={fused}: {e0.elapsed_time(e1) / 20:.3f} ms/step  apply={k['k512_apply']:.3f} normalise={k['k_peak_normalise']:.3f}", flush=True)
    outs[fused] = eng.out.clone()
    del eng
print("bit-identical:", bool(torch.equal(outs[False], outs[True])))

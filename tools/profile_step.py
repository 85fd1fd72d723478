"""Short single-GPU run of the fused oracle-mask MVDR step for ncu (keep it small: ncu replays every kernel).
usage: python tools/profile_step.py [B] [steps] [fused|separate]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import avzoom  # noqa: E402
from avzoom import pipeline, synth  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 296
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
cfg = avzoom.PRESETS["baseline_oracle"]
mix, tgt, itf = synth.make_batch(2, 8, 4.0, 3)
reps = (B + 7) // 8
mix = torch.from_numpy(mix).cuda().repeat(reps, 1, 1)[:B].contiguous()
tgt = torch.from_numpy(tgt).cuda().repeat(reps, 1)[:B].contiguous()
itf = torch.from_numpy(itf).cuda().repeat(reps, 1)[:B].contiguous()
fused = (sys.argv[3] if len(sys.argv) > 3 else "fused") == "fused"
eng = pipeline.OracleMvdr(cfg, B, mix.shape[-1], mix.device, fused=fused)
for _ in range(steps):
    eng.run(mix, tgt, itf)
torch.cuda.synchronize()
print("ok", B, steps, float(eng.out.abs().max()))

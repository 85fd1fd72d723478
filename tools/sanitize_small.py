"""Small end-to-end run for compute-sanitizer (memcheck / racecheck): fast path (kept + recompute), generic path,
streaming step, features - on tiny shapes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, avzoom
from avzoom import synth, pipeline, stream
cfg = avzoom.PRESETS["baseline_oracle"]
mix, tgt, itf = synth.make_batch(1, 3, 0.3, 2)
m, t, i = (torch.from_numpy(a).cuda() for a in (mix, tgt, itf))
for keep in (True, False):
    e = pipeline.OracleMvdr(cfg, 3, mix.shape[-1], m.device, keep_spectrum=keep)
    e.run(m, t, i)
avzoom.oracle_mask_mvdr(m, t, i, avzoom.PRESETS["oracle_debug"])
X = avzoom.wave_features(m, 512, 128)
Y = avzoom.stft(m, 1024, 512)
avzoom.istft(Y, 1024, 512)
mask = torch.rand((3, 257, avzoom.num_frames(mix.shape[-1], 512, 128)), device="cuda")
avzoom.learned_mask_mvdr(m, mask, avzoom.PRESETS["baseline_learned"])
s = stream.MvdrStream(5, cfg)
for h in range(6):
    s.step(torch.randn((5, 2, 128), device="cuda") * 0.1, torch.rand((5, 257), device="cuda"))
avzoom.sir_scores(m[:, 0, :4736].contiguous(), t[:, :4736].contiguous(), i[:, :4736].contiguous())
torch.cuda.synchronize()
print("sanitize run ok")

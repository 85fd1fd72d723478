"""Hot-path kernels of config 3 without the U-Net (a random mask stands in): for ncu launch lists.
usage: python tools/learned_kernels_probe.py [B]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import avzoom  # noqa: E402
from avzoom import synth  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
cfg = avzoom.PRESETS["baseline_learned"]
mix8, _, _ = synth.make_batch(3, 8, 4.0, 3)
mix = torch.from_numpy(mix8).cuda().repeat((B + 7) // 8, 1, 1)[:B].contiguous()
torch.manual_seed(0)
T = avzoom.num_frames(mix.shape[-1], cfg.n_fft, cfg.hop)
mask = torch.rand((B, cfg.n_freq, T), device="cuda")
for _ in range(3):
    X = avzoom.wave_features(mix, cfg.n_fft, cfg.hop)
    out = avzoom.learned_mask_mvdr(mix, mask, cfg)
torch.cuda.synchronize()
print("ok", list(X.shape), list(out.shape))

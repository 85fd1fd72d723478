"""One launch of each n_fft 1024 / hop 512 fast-path kernel after a warm-up, for ncu:
ncu --set full --clock-control none --import-source on -k regex:"k1024_|k_cov_finalize|k_mvdr_weights" -s 9 -c 9 -o gpurun_out/prof python tools/profile_1024.py [B]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import avzoom  # noqa: E402
from avzoom import ops, synth  # noqa: E402

B, L = (int(sys.argv[1]) if len(sys.argv) > 1 else 512), 32000
cfg = avzoom.PRESETS["full_audio"]
mix8, _, _ = synth.make_batch(3, 8, 2.0, 3)
mix = torch.from_numpy(mix8).cuda().repeat(B // 8, 1, 1).contiguous()
T = avzoom.num_frames(L, cfg.n_fft, cfg.hop)
mask = torch.rand((B, cfg.n_freq, T), device="cuda")
spec = ops.alloc_kept_spectrum(mix, cfg)
for _ in range(2):   # the first round warms up, the second is the profiled one
    X = avzoom.wave_features(mix, cfg.n_fft, cfg.hop)
    for sp in (None, spec):   # recomputing and kept-spectrum variants of both passes
        Rp, _ = ops.wave_masked_covariance(mix, mask, cfg, sp)
        w = ops.mvdr_weights(Rp, ops.steering_vectors(cfg, mix.device), cfg)
        out, peak = ops.mvdr_apply(mix, w, cfg, mask=mask, spec=sp)
    torch.cuda.synchronize()
print("ok", tuple(X.shape), tuple(out.shape))

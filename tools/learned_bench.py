"""BASELINE config 3: log-mag + IPD features -> random-init U-Net mask predictor -> mask-MVDR, batch x 4 s at the C2 STFT
shape (n_fft 512 / hop 128).  Reports the hot-path stages (this project's kernels) and the U-Net (torch / cuDNN) apart.
usage: python tools/learned_bench.py [B] [iters]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import avzoom  # noqa: E402
from avzoom import synth  # noqa: E402
from avzoom.core import models  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 10   # the first iterations pay for the 264 MB output buffers
cfg = avzoom.PRESETS["baseline_learned"]
mix8, _, _ = synth.make_batch(3, 8, 4.0, 3)
mix = torch.from_numpy(mix8).cuda().repeat((B + 7) // 8, 1, 1)[:B].contiguous()
torch.manual_seed(0)
net = models.FreqPreservingUNet().eval().cuda()


def timed(fn):
    fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        r = fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters, r


t_feat, X = timed(lambda: avzoom.wave_features(mix, cfg.n_fft, cfg.hop))
sub = 16                                   # U-Net activations at 257 x 501 are large: run it in sub-batches


def unet():
    with torch.no_grad():
        return torch.cat([net(X[i:i + sub]) for i in range(0, B, sub)]).float().contiguous()


t_net, mask = timed(unet)
t_mvdr, out = timed(lambda: avzoom.learned_mask_mvdr(mix, mask, cfg))
audio_s = B * 4.0
print(json.dumps({"config": "BASELINE config 3: features -> FreqPreservingUNet (random init, eval) -> learned-mask MVDR",
                  "batch": B, "features_ms": t_feat, "unet_ms": t_net, "mask_mvdr_ms": t_mvdr,
                  "hot_path_audio_s_per_s": audio_s / ((t_feat + t_mvdr) * 1e-3),
                  "end_to_end_audio_s_per_s": audio_s / ((t_feat + t_net + t_mvdr) * 1e-3),
                  "out_shape": list(out.shape), "finite": bool(torch.isfinite(out).all())}))

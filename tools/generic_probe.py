"""Timing of the n_fft 1024 / hop 512 kernels the reference's learned pipelines use: 512 windows of 2 s.
python tools/generic_probe.py                      register-resident fast path (avz_opt1024.cu)
AVZ_FORCE_GENERIC=1 python tools/generic_probe.py  generic shared-memory FFT kernels (A/B)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import avzoom  # noqa: E402
from avzoom import synth  # noqa: E402

B, L = 512, 32000
cfg = avzoom.PRESETS["full_audio"]
mix8, _, _ = synth.make_batch(3, 8, 2.0, 3)
mix = torch.from_numpy(mix8).cuda().repeat(B // 8, 1, 1).contiguous()
T = avzoom.num_frames(L, cfg.n_fft, cfg.hop)
mask = torch.rand((B, cfg.n_freq, T), device="cuda")


def timed(fn, iters=5):
    fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


from avzoom import ops  # noqa: E402
Rp, _ = ops.wave_masked_covariance(mix, mask, cfg, None)
wts = ops.mvdr_weights(Rp, ops.steering_vectors(cfg, mix.device), cfg)
t_cov = timed(lambda: ops.wave_masked_covariance(mix, mask, cfg, None), 20)
t_apply = timed(lambda: ops.mvdr_apply(mix, wts, cfg, mask=mask), 20)
spec = ops.alloc_kept_spectrum(mix, cfg)
if spec is not None:
    t_cov_k = timed(lambda: ops.wave_masked_covariance(mix, mask, cfg, spec), 20)
    t_apply_k = timed(lambda: ops.mvdr_apply(mix, wts, cfg, mask=mask, spec=spec), 20)
    print(f"  kept spectrum: pass A {t_cov_k:.3f} ms, pass B {t_apply_k:.3f} ms")
import dataclasses  # noqa: E402
t_apply_nomask = timed(lambda: ops.mvdr_apply(mix, wts, dataclasses.replace(cfg, post="none")), 20)
print(f"  pass B without the post-filter mask reads {t_apply_nomask:.3f} ms")
print(f"  pass A (covariance + finalize) {t_cov:.3f} ms, pass B (apply) {t_apply:.3f} ms")
t_feat = timed(lambda: avzoom.wave_features(mix, cfg.n_fft, cfg.hop))
t_mvdr = timed(lambda: avzoom.learned_mask_mvdr(mix, mask, cfg))
audio_s = B * L / 16000.0
which = "generic" if os.environ.get("AVZ_FORCE_GENERIC") == "1" else "fast"
print(f"1024/512 {which} path, {B} x 2 s: features {t_feat:.3f} ms, learned-mask MVDR {t_mvdr:.3f} ms "
      f"-> {audio_s / ((t_feat + t_mvdr) * 1e-3) / 1e6:.3f} M audio-s/s")

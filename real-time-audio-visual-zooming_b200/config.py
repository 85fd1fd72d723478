"""One frozen record of every constant in which the reference's call sites differ
(SURVEY.md 8-A2), with a preset named after each site."""
from __future__ import annotations

import dataclasses
from typing import Optional

import numpy as np

from . import _lib


@dataclasses.dataclass(frozen=True)
class MvdrConfig:
    fs: float = 16000.0
    n_fft: int = 512
    hop: int = 128
    mic_dist: float = 0.01          # masked_mvdr.py:10 (D)
    c: float = 343.0                # masked_mvdr.py:11
    angle_deg: float = 90.0         # oracle_debug.py:23
    sigma: float = 1.0              # oracle_debug.py:24
    hp_hz: Optional[float] = 100.0  # oracle_debug.py:69; None = no high-pass (batch_mvdr)
    hp_mode: str = "zero"           # 'zero' (oracle_debug.py:69) | 'mic0' (Final_pipeline/src/inference.py:51-53)
    sqrt_eps: float = 0.0           # tf_lite_version/inference.py:111 uses 1e-10
    norm_eps: float = 1e-6          # oracle_debug.py:64
    w_eps: float = 1e-10            # oracle_debug.py:77
    post: str = "one_minus_noise"   # one_minus_noise | floor | mask | none
    post_floor: float = 0.05        # full_audio.../inference.py:116
    peak_eps: Optional[float] = 0.0  # oracle_debug.py:94 (0); masked_mvdr.py:128 (1e-6); None = no peak norm

    @property
    def n_freq(self) -> int:
        return self.n_fft // 2 + 1

    def freqs(self) -> np.ndarray:
        return np.fft.rfftfreq(self.n_fft, 1.0 / self.fs)

    def hp_bins(self) -> int:
        if self.hp_hz is None:
            return 0
        return int(np.sum(self.freqs() < self.hp_hz))

    def to_c(self) -> "_lib.AvzMvdrCfg":
        post = {"none": _lib.POST_NONE, "one_minus_noise": _lib.POST_ONE_MINUS_NOISE,
                "floor": _lib.POST_FLOOR, "mask": _lib.POST_MASK}[self.post]
        if self.hp_hz is None:
            hp = _lib.HP_NONE
        else:
            hp = {"zero": _lib.HP_ZERO, "mic0": _lib.HP_MIC0}[self.hp_mode]
        return _lib.AvzMvdrCfg(self.sigma, self.norm_eps, self.sqrt_eps, self.w_eps, self.hp_bins(), hp, post,
                               self.post_floor)


PRESETS = {
    # BASELINE C1/C2: oracle_debug.py arithmetic at n_fft 512 / hop 128 (nb cell6:31-32)
    "baseline_oracle": MvdrConfig(),
    # rt_av_zoom/core/oracle_debug.py as written (N_HOP = 256 passed as noverlap)
    "oracle_debug": MvdrConfig(hop=256),
    # rt_av_zoom/core/masked_mvdr.py:9-18,76-128
    "masked_mvdr": MvdrConfig(hop=256, sigma=1e-7, post="none", peak_eps=1e-6),
    # rt_av_zoom/core/full_audio_generating_pipeline/inference.py with its config.json
    "full_audio": MvdrConfig(n_fft=1024, hop=512, mic_dist=0.04, sigma=1e-5, post="floor", peak_eps=None),
    # rt_av_zoom/core/tf_lite_version/inference.py (batch_mvdr: sqrt eps, no high-pass)
    "tf_lite": MvdrConfig(n_fft=1024, hop=512, mic_dist=0.04, sigma=1e-5, hp_hz=None, sqrt_eps=1e-10, post="floor",
                          peak_eps=1e-9),
    # BASELINE C3: learned mask at the C2 STFT shape
    "baseline_learned": MvdrConfig(mic_dist=0.04, sigma=1e-5, post="floor", peak_eps=None),
}

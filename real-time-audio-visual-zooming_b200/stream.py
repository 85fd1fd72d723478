"""Streaming mask-MVDR (BASELINE config 4): recursive exponentially-smoothed covariance, one hop
of 128 samples per call across many concurrent 2-mic streams.  The reference has no streaming mode; the recursion is
defined here (see include/avzoom.h) and checked against `oracle.streaming_mvdr`: parity unpinned by construction."""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import _lib
from .config import MvdrConfig, PRESETS
from .ops import _ptr, _stream, steering_vectors

INT_MAX = 2 ** 31 - 1


class MvdrStream:
    """State of `n_streams` independent streams (n_fft 512, hop 128)."""

    LATENCY_HOPS = 3   # call h returns output samples [(h-3)*128, (h-2)*128)

    def __init__(self, n_streams: int, cfg: MvdrConfig = PRESETS["baseline_oracle"], lam: float = 0.95, device="cuda"):
        if cfg.n_fft != 512 or cfg.hop != 128:
            raise ValueError("streaming mode is built for n_fft 512 / hop 128")
        self.cfg, self.lam, self.S = cfg, float(lam), int(n_streams)
        self.lib = _lib.load()
        nbytes = self.lib.avz_stream_state_bytes(self.S)
        self.state = torch.zeros((nbytes // 4,), dtype=torch.float32, device=device)
        self.d = steering_vectors(cfg, self.state.device)
        self.cc = cfg.to_c()
        self.h = 0
        self.out = torch.empty((self.S, 128), dtype=torch.float32, device=device)
        _lib.check(self.lib.avz_init(512), "avz_init")

    def reset(self) -> None:
        """Forget everything: the next step() is hop 0 of a new set of streams."""
        self.h = 0
        self.state.zero_()

    def step(self, hop_in: torch.Tensor, noise_w: Optional[torch.Tensor] = None, t_end: int = INT_MAX) -> torch.Tensor:
        """hop_in [S,2,128] f32 (CUDA), noise_w [S,257] or None -> [S,128] output hop (valid from the 4th call on)."""
        _lib.check(self.lib.avz_stream_step_f32(_ptr(self.state), _ptr(hop_in), _ptr(noise_w), _ptr(self.d), self.S,
                                                self.h - 1, int(t_end), self.lam, C.byref(self.cc), _ptr(self.out),
                                                _stream()), "avz_stream_step_f32")
        self.h += 1
        return self.out

    def run(self, mix: torch.Tensor, noise_w: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Whole recordings through the streaming path: mix [S,2,L] (L a multiple of 128), noise_w [S,257,T] per-frame
        noise weights (T = L/128 + 1) -> [S, L]; equals `oracle.streaming_mvdr` on each stream."""
        S, _, L = mix.shape
        assert L % 128 == 0 and S == self.S
        self.reset()                      # a recording starts from silence: frame index 0, empty covariance and tails
        H = L // 128
        T = H + 1
        out = torch.empty((S, L), dtype=torch.float32, device=mix.device)
        zero = torch.zeros((S, 2, 128), dtype=torch.float32, device=mix.device)
        for h in range(H + 3):
            x = mix[:, :, h * 128:(h + 1) * 128].contiguous() if h < H else zero
            t = h - 1
            m = noise_w[:, :, t].contiguous() if (noise_w is not None and 0 <= t < T) else None
            y = self.step(x, m, t_end=T)
            if h >= 3:
                out[:, (h - 3) * 128:(h - 2) * 128] = y
        return out

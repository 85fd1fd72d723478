"""avzoom: B200-native (sm_100a) mask-driven 2-mic MVDR beamforming.

Drop-in for the hot path of Senpai-sama06/real-time-audio-visual-zooming: the public names below keep the
reference's signatures and array layouts; the arithmetic runs in hand-written CUDA kernels behind the C ABI of
`libavzoom.so` (include/avzoom.h).  Import is light (no CUDA work); the first op call loads the library and
raises if it is missing - there is no CPU fallback.
"""
from .config import MvdrConfig, PRESETS  # noqa: F401
from .ops import *  # noqa: F401,F403
from . import ops as _ops

__all__ = ["MvdrConfig", "PRESETS"] + list(_ops.__all__)
__version__ = "0.1.0"

"""Mirror of rt_av_zoom/core/tf_lite_version/world_building.py:39-110 (the anechoic far-field mixer) on the GPU.

Same names and argument meaning as the reference; audio arrays may be numpy (numpy comes back) or CUDA tensors.
File I/O (`load_resample`, LJ Speech download) is out of scope: `mix_sources` takes the raw sources directly.
"""
from __future__ import annotations

import numpy as np

from .. import ops

FS = 16000
D = 0.04
C = 343.0
ANGLE_TARGET = 90.0
ANGLE_INTERFERER_A = 40.0
ANGLE_INTERFERER_B = 130.0


def calculate_far_field_delays(azimuth_deg, d, c):
    """world_building.py:40-44 -> (tau_m1, tau_m2) seconds."""
    theta_rad = np.deg2rad(azimuth_deg)
    return (d / 2) * np.cos(theta_rad - 0) / c, (d / 2) * np.cos(theta_rad - np.pi) / c


def apply_frac_delay(y, delay_sec, fs):
    """world_building.py:46-52."""
    return ops.fractional_delay(y, float(delay_sec), float(fs))


def mix_sources(sources, angles_deg=(ANGLE_TARGET, ANGLE_INTERFERER_A, ANGLE_INTERFERER_B), d=D, c=C, fs=FS):
    """world_building.py:61-93 (`mix_and_save` up to the sf.write calls): sources (S,L) or (B,S,L), source 0 is the
    target.  -> (mix (..,2,L), tgt_ref (..,L), int_ref (..,L)), all divided by max|mix| + 1e-9."""
    n_src = sources.shape[-2]
    if len(angles_deg) < n_src:
        raise ValueError(f"{n_src} sources but {len(angles_deg)} angles")
    delays = [calculate_far_field_delays(a, d, c) for a in angles_deg[:n_src]]
    return ops.far_field_mix(sources, delays, fs, peak_eps=1e-9)

"""Stand-in for Final_pipeline/src/simulation.py:58 `generate_scene`.  The reference simulates a reverberant room
with pyroomacoustics on LJ Speech (neither is available offline; SURVEY.md section 2 row 10 marks it out of scope), so
scenes here are the anechoic far-field mixtures of `synth` written in the same on-disk layout:
{SIM_DIR}/{run_name}/mixture.wav, target_reference.wav, interference_reference.wav (16 kHz PCM16)."""
from __future__ import annotations

import os
import zlib

from .. import synth, wavio
from . import config


def generate_scene(run_name, dataset="synthetic", reverb=False, n_interferers=2, snr_target=50, duration_s=4.0):
    out_dir = os.path.join(config.SIM_DIR, run_name)
    os.makedirs(out_dir, exist_ok=True)
    seed = zlib.crc32(run_name.encode()) & 0x7FFFFFFF
    mix, tgt, itf = synth.make_mixture(seed, int(duration_s * config.FS), n_interferers, d=config.MIC_DIST,
                                       c=config.C_SPEED, fs=config.FS)
    mix_path = os.path.join(out_dir, "mixture.wav")
    wavio.write(mix_path, mix.T, config.FS)
    wavio.write(os.path.join(out_dir, "target_reference.wav"), tgt, config.FS)
    wavio.write(os.path.join(out_dir, "interference_reference.wav"), itf, config.FS)
    print(f"[SIM] Scene '{run_name}' written to {out_dir}")
    return mix_path

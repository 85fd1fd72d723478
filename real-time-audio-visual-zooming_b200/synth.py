"""Synthetic far-field 2-mic mixtures for benchmarks and tests (LJ Speech is not available offline).

The mixing physics restates the reference's world builder
(rt_av_zoom/core/tf_lite_version/world_building.py:40-93): per-source far-field delays
tau = +-(d/2) cos(theta)/c, whole-signal rFFT phase-ramp fractional delay, target at 90 degrees and
interferers at 40/130/65/155/20... degrees, references = the mic-1 images, everything divided by
max|mix| + 1e-9.  The *sources* are speech-like noise (band-limited components under syllabic envelopes)
so that the ideal binary mask is non-trivial; recipe and seeding follow SURVEY.md 8-D:
utterance u of config c uses numpy.random.default_rng(1_000_003 * c + u).
Data are kept as float32 in memory (no PCM16 round trip).
"""
from __future__ import annotations

import numpy as np
import scipy.signal

INTERFERER_ANGLES = (40.0, 130.0, 65.0, 155.0, 20.0, 110.0, 75.0, 170.0)
TARGET_ANGLE = 90.0


def far_field_delays(az_deg: float, d: float, c: float):
    """world_building.py:40-44."""
    th = np.deg2rad(az_deg)
    return (d / 2) * np.cos(th - 0) / c, (d / 2) * np.cos(th - np.pi) / c


def _delay_pair(y: np.ndarray, t1: float, t2: float, fs: float):
    """world_building.py:46-52 applied for both mics with one forward transform."""
    n = len(y)
    spec = np.fft.rfft(y)
    fr = np.fft.rfftfreq(n, 1.0 / fs)
    ramp = -2j * np.pi * fr
    return np.fft.irfft(spec * np.exp(ramp * t1), n=n), np.fft.irfft(spec * np.exp(ramp * t2), n=n)


def speech_like(rng: np.random.Generator, n: int, fs: float = 16000.0) -> np.ndarray:
    """Sum of three band-limited noise components, each under its own syllabic envelope."""
    out = np.zeros(n)
    for _ in range(3):
        fc = rng.uniform(300.0, 3200.0)
        bw = rng.uniform(300.0, 900.0)
        lo, hi = max(120.0, fc - bw / 2), min(0.45 * fs, fc + bw / 2)
        b, a = scipy.signal.butter(2, [lo, hi], btype="band", fs=fs)
        carrier = scipy.signal.lfilter(b, a, rng.standard_normal(n))
        onsets = (rng.random(n) < 5.0 / fs) * rng.uniform(0.3, 1.0, n)
        pole = np.exp(-1.0 / (0.06 * fs))
        env = scipy.signal.lfilter([1.0 - pole], [1.0, -pole], onsets)
        env /= env.max() + 1e-12
        out += carrier * env
    return out / (np.max(np.abs(out)) + 1e-12)


def make_mixture(seed: int, n_samples: int, n_interferers: int, d: float = 0.04, c: float = 343.0,
                 fs: float = 16000.0):
    """One utterance -> (mix (2,L), tgt (L,), itf (L,)) float32  (world_building.py:61-93 semantics)."""
    rng = np.random.default_rng(seed)
    angles = (TARGET_ANGLE,) + tuple(INTERFERER_ANGLES[i % len(INTERFERER_ANGLES)] for i in range(n_interferers))
    m1 = np.zeros(n_samples)
    m2 = np.zeros(n_samples)
    tgt = np.zeros(n_samples)
    itf = np.zeros(n_samples)
    for idx, ang in enumerate(angles):
        src = speech_like(rng, n_samples, fs)
        t1, t2 = far_field_delays(ang, d, c)
        s1, s2 = _delay_pair(src, t1, t2, fs)
        m1 += s1
        m2 += s2
        if idx == 0:
            tgt += s1
        else:
            itf += s1
    mix = np.stack([m1, m2])
    norm = np.max(np.abs(mix)) + 1e-9
    return (mix / norm).astype(np.float32), (tgt / norm).astype(np.float32), (itf / norm).astype(np.float32)


def make_batch(config_id: int, n_utt: int, dur_s: float, n_interferers: int, start: int = 0, fs: float = 16000.0,
               workers: int = 0):
    """Utterances start..start+n_utt-1 of a BASELINE config -> mix (B,2,L), tgt (B,L), itf (B,L) float32."""
    L = int(round(dur_s * fs))
    seeds = [1_000_003 * config_id + start + u for u in range(n_utt)]
    args = [(s, L, n_interferers) for s in seeds]
    if workers and n_utt >= 4 * workers:
        import multiprocessing as mp
        with mp.get_context("fork").Pool(workers) as pool:
            res = pool.starmap(make_mixture, args, chunksize=max(1, n_utt // (8 * workers)))
    else:
        res = [make_mixture(*a) for a in args]
    mix = np.stack([r[0] for r in res])
    tgt = np.stack([r[1] for r in res])
    itf = np.stack([r[2] for r in res])
    return mix, tgt, itf

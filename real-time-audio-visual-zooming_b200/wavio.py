"""16 kHz PCM16 WAV in / out through the stdlib `wave` module (the reference uses `soundfile`, absent here).

`read` mirrors `soundfile.read(path, dtype='float32')`: int16 / 32768 as float32, shape (n,) or (n, channels).
`write` mirrors `soundfile.write(path, data, fs)` for the default PCM_16 subtype: libsndfile scales floats by
0x7FFF and rounds to nearest (values outside [-1, 1] are clipped here).
"""
from __future__ import annotations

import wave

import numpy as np


def read(path: str, dtype: str = "float32"):
    with wave.open(path, "rb") as w:
        n, ch, fs, sw = w.getnframes(), w.getnchannels(), w.getframerate(), w.getsampwidth()
        raw = w.readframes(n)
    if sw != 2:
        raise ValueError(f"{path}: only PCM16 WAV is supported (sample width {sw})")
    data = np.frombuffer(raw, dtype="<i2").astype(np.float64) / 32768.0
    if ch > 1:
        data = data.reshape(n, ch)
    return data.astype(dtype), fs


def write(path: str, data, fs: int) -> None:
    data = np.asarray(data, dtype=np.float64)
    pcm = np.clip(np.rint(data * 32767.0), -32768, 32767).astype("<i2")
    ch = 1 if pcm.ndim == 1 else pcm.shape[1]
    with wave.open(path, "wb") as w:
        w.setnchannels(ch)
        w.setsampwidth(2)
        w.setframerate(int(fs))
        w.writeframes(np.ascontiguousarray(pcm).tobytes())

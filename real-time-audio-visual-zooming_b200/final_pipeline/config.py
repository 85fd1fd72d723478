"""Constants of the final pipeline under the names the reference's drivers import (Final_pipeline/src/config.py).

Values are the reference's; they are grouped here by what consumes them.  Directories hang off AVZOOM_PROJECT_ROOT
(default: ./Final_pipeline under the current directory) instead of the location of this file, because this package is
installed read-only next to the CUDA library."""
import os as _os


def _under(root: str, *parts: str) -> str:
    return _os.path.join(root, *parts)


PROJECT_ROOT = _os.environ.get("AVZOOM_PROJECT_ROOT") or _under(_os.getcwd(), "Final_pipeline")
DATA_DIR = _under(PROJECT_ROOT, "data")
RAW_DATA_DIR, SIM_DIR, RESULTS_DIR = (_under(DATA_DIR, leaf) for leaf in ("raw", "simulated", "results"))

# STFT / chunking of the learned pipeline: 2 s windows of 16 kHz audio, 1024-point frames at 50 % overlap
FS = 16_000
WIN_SIZE = 2 * FS
N_FFT = 1 << 10
HOP_LEN = N_FFT >> 1

# acoustics: speed of sound and the two-microphone array (8 cm apart, centred in a 4.9 m cube at 1.5 m height)
C_SPEED = 343.0
MIC_DIST = 0.08
ROOM_DIM = [4.9] * 3
MIC_LOCS_SIM = [[round(ROOM_DIM[0] / 2 + s * MIC_DIST / 2, 2), round(ROOM_DIM[1] / 2, 2), 1.5] for s in (-1, 1)]

# defaults of the (out-of-scope) room simulator, kept so that callers' imports resolve
RT60_TARGET = 0.5
SIR_TARGET_DB = 0

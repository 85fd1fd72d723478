"""Final_pipeline/src/config.py:1-28 (paths relative to this package's data directory)."""
import os

PROJECT_ROOT = os.environ.get("AVZOOM_PROJECT_ROOT", os.path.join(os.getcwd(), "Final_pipeline"))
DATA_DIR = os.path.join(PROJECT_ROOT, "data")
RAW_DATA_DIR = os.path.join(DATA_DIR, "raw")
SIM_DIR = os.path.join(DATA_DIR, "simulated")
RESULTS_DIR = os.path.join(DATA_DIR, "results")

FS = 16000
C_SPEED = 343.0
N_FFT = 1024
HOP_LEN = 512
WIN_SIZE = 32000

ROOM_DIM = [4.9, 4.9, 4.9]
RT60_TARGET = 0.5
SIR_TARGET_DB = 0

MIC_LOCS_SIM = [[2.41, 2.45, 1.5], [2.49, 2.45, 1.5]]
MIC_DIST = 0.08

"""Mirror of Final_pipeline/src/inference.py: `get_steering_vector_single` (:16-26), `hybrid_hard_null_bf` (:28-98)
and `enhance_audio(run_name, input_path, model_path)` (:144-238).

The chunk loop (2 s windows, 50 % overlap, count-averaged OLA clipped to the input length, peak normalisation with
1e-9), the hybrid hard-null beamformer (interference covariance -> principal eigenvector -> 2x2 constraint solve with
a condition-number fallback to delay-and-sum, mic-0 pass-through below 200 Hz) and the post-filter S * mask are the
reference's; all windows of a recording run as one batch through the fused kernels.  The mask estimator is a torch
module ((B,2,F,T) log-mag + IPD features -> (B,F,T)) because the reference's `.tflite` blob is not shipped
(.MISSING_LARGE_BLOBS); its 4-channel "physics" feature layout is available as `ops.wave_features(..., 'physics')`."""
from __future__ import annotations

import os
import time

import numpy as np
import torch

from .. import ops, wavio
from ..config import MvdrConfig
from ..core import chunked
from . import config

ANGLE_TARGET = 90.0
FREQ_BINS = (config.N_FFT // 2) + 1
N_MICS = 2

# constants of Final_pipeline/src/inference.py as one record: no diagonal loading / MVDR here, the weights come from
# the hybrid null solve; pass B only needs the STFT shape and the post-filter (S * mask, :219)
FINAL_CFG = MvdrConfig(fs=config.FS, n_fft=config.N_FFT, hop=config.HOP_LEN, mic_dist=config.MIC_DIST, c=config.C_SPEED,
                       angle_deg=ANGLE_TARGET, sigma=0.0, hp_hz=200.0, hp_mode="mic0", post="mask", peak_eps=None)


def get_steering_vector_single(f, angle_deg, d, c):
    """Steering vector of one bin, phase-normalised to mic 0 (Final_pipeline/src/inference.py:16-26) -> (2, 1)."""
    theta = np.deg2rad(angle_deg)
    tau1 = (d / 2) * np.cos(theta) / c
    tau2 = (d / 2) * np.cos(theta - np.pi) / c
    omega = 2 * np.pi * f
    v = np.array([[np.exp(-1j * omega * tau1)], [np.exp(-1j * omega * tau2)]])
    return v / (v[0] + 1e-10)


def _steering(f_bins, device, wide=False):
    f = np.asarray(f_bins, dtype=np.float64)
    th = np.deg2rad(ANGLE_TARGET)
    tau1 = (config.MIC_DIST / 2) * np.cos(th) / config.C_SPEED
    tau2 = (config.MIC_DIST / 2) * np.cos(th - np.pi) / config.C_SPEED
    om = 2 * np.pi * f
    d = np.stack([np.exp(-1j * om * tau1), np.exp(-1j * om * tau2)], axis=1)
    return torch.from_numpy(d.astype(np.complex128 if wide else np.complex64)).to(device)


def hybrid_hard_null_bf(Y, mask, f_bins, degenerate="raise"):
    """Y (M=2, F, T) complex STFT, mask (F, T) target probabilities, f_bins (F,) Hz -> beamformed (F, T) complex.
    numpy in -> numpy out; CUDA tensors in -> CUDA tensor out.  The arithmetic follows the dtype of Y like numpy:
    complex128 (what scipy.signal.stft hands the reference) runs on the float64 operators - the eigenvector, the
    condition-number threshold and the constraint solve amplify float32 rounding beyond the 1e-4 parity budget.

    `degenerate`: what a bin above the bypass does when its interference covariance has no principal direction with a
    mic-0 component (e.g. mask == 1 in every frame of the bin).  The reference divides by zero there, its constraint
    matrix is NaN and np.linalg.cond raises LinAlgError("SVD did not converge") out of the function
    (Final_pipeline/src/inference.py:66-81; pinned in tests/golden/ref_chunk_drivers.npz): 'raise' does the same,
    'nan' returns NaN in that bin, 'das' falls back to delay-and-sum (what the batched chunk path does)."""
    if degenerate not in ("raise", "nan", "das"):
        raise ValueError("degenerate must be 'raise', 'nan' or 'das'")
    is_np = isinstance(Y, np.ndarray)
    Yt = torch.as_tensor(Y).cuda() if is_np else Y
    wide = Yt.dtype == torch.complex128
    if not wide:
        Yt = Yt.to(torch.complex64)
    mt = torch.as_tensor(mask).to(Yt.device)
    m_int = (1.0 - mt).to(torch.float64 if wide else torch.float32)      # `mask_int = 1.0 - mask` in the mask's own dtype
    f = np.asarray(f_bins, dtype=np.float64)
    R = ops.masked_covariance(Yt, m_int, sqrt_eps=0.0, norm_eps=1e-6, packed=True)
    w = ops.hybrid_null_weights(R, _steering(f, Yt.device, wide), int(np.sum(f < 200)), zero_cov_nan=(degenerate != "das"))
    if degenerate == "raise" and bool(torch.isnan(w.real).any()):
        raise np.linalg.LinAlgError("SVD did not converge")
    S = ops.beamform(w, Yt)
    return S.cpu().numpy() if is_np else S


class TFLiteBeamformer:
    """Name and call contract of the reference's physics-aware wrapper (Final_pipeline/src/inference.py:102-141):
    `TFLiteBeamformer(model_path).predict_mask(log_mag (F,T), raw_ipd (F,T)) -> (F,T)`.

    The TFLite interpreter is replaced by a torch module: `model_path` is a TorchScript file (`torch.jit.load`), or
    pass `model=` directly; it receives the reference's input tensor, (1, F, T, 4) float32 NHWC =
    [log_mag, sin ipd, cos ipd, linspace(0,1,F) tiled over T], and returns the mask in any shape that squeezes to
    (F, T).  A missing file raises FileNotFoundError like the reference.  (Batches of windows take the fused kernel
    instead: `ops.wave_features(chunks, n_fft, hop, 'physics')` builds the same tensor straight from the waveform.)"""

    def __init__(self, model_path, model=None):
        if model is None:
            if not os.path.exists(model_path):
                raise FileNotFoundError(f"Model file not found: {model_path}")
            model = torch.jit.load(model_path, map_location="cuda" if torch.cuda.is_available() else "cpu")
        self.model = model
        self.freq_map = np.linspace(0, 1, FREQ_BINS, dtype=np.float32)[:, np.newaxis]

    def predict_mask(self, log_mag, raw_ipd):
        is_np = isinstance(log_mag, np.ndarray)
        lm = torch.as_tensor(log_mag, dtype=torch.float32)
        ipd = torch.as_tensor(raw_ipd, dtype=torch.float32).to(lm.device)
        fmap = torch.linspace(0, 1, lm.shape[0], dtype=torch.float32, device=lm.device)[:, None].expand(-1, lm.shape[1])
        x = torch.stack([lm, torch.sin(ipd), torch.cos(ipd), fmap], dim=-1)[None]
        with torch.no_grad():
            out = self.model(x).float().squeeze()
        return out.cpu().numpy() if is_np else out


def enhance_chunks(chunks: torch.Tensor, model, cfg: MvdrConfig = FINAL_CFG) -> torch.Tensor:
    """One batch of 2 s windows (n, 2, win) -> (n, iSTFT length): features (STFT fused) -> mask model -> interference
    covariance (STFT fused) -> hybrid-null weights -> beamform + S * mask + iSTFT (fused)."""
    X = ops.wave_features(chunks, cfg.n_fft, cfg.hop, "logmag_ipd")
    with torch.no_grad():
        mask = model(X).float().contiguous()
    # interference covariance and weights in float64 (the null solve is ill-conditioned); pass B stays on the fused
    # float32 kernels: rounding w to complex64 moves the output by ~1e-7
    Rp, _ = ops.wave_masked_covariance(chunks, mask, cfg, wide=True)
    w = ops.hybrid_null_weights(Rp, _steering(cfg.freqs(), chunks.device, wide=True), cfg.hp_bins(), round_to_f32=True)
    out, _ = ops.mvdr_apply(chunks, w, cfg, mask=mask)
    return out


def enhance_audio(run_name, input_path, model_path, model=None):
    """Reads `input_path`, writes {RESULTS_DIR}/{run_name}_results/{run_name}_enhanced.wav."""
    result_dir = os.path.join(config.RESULTS_DIR, f"{run_name}_results")
    os.makedirs(result_dir, exist_ok=True)
    output_path = os.path.join(result_dir, f"{run_name}_enhanced.wav")
    print(f"[INF] Processing {input_path}")
    print(f"[INF] Saving to  {output_path}")
    y, sr = wavio.read(input_path, dtype="float32")
    if sr != config.FS:
        print(f"Warning: SR mismatch. Input: {sr}, Config: {config.FS}")
    if y.ndim == 1:
        print("Error: Input is mono. Requires 2 channels.")
        return
    if model is None:
        try:
            model = chunked.load_mask_model(model_path if model_path and os.path.exists(model_path) else None)
        except Exception as e:  # same contract as the reference: report and give up on this run
            print(f"Failed to load mask model: {e}")
            return
    start_time = time.time()
    enh = chunked.ChunkedEnhancer(FINAL_CFG, config.WIN_SIZE, clip_to_input=True, weights="hybrid_null",
                                  final_peak_eps=1e-9, steering=lambda dev: _steering(FINAL_CFG.freqs(), dev, wide=True))
    final = enh(chunked.to_planar(y), model)[0].cpu().numpy()
    print(f"Total processing time: {time.time() - start_time:.3f}s")
    wavio.write(output_path, final, config.FS)
    return output_path

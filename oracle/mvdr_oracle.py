"""Float64 numpy/scipy restatement of the reference's mask-driven MVDR path.

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).  Every function cites the
reference lines (relative to ``/root/reference``) whose arithmetic it follows.
All arrays are float64 / complex128: inputs are cast on entry, exactly as the
parity plan in BASELINE.md section 2 asks ("the reference's own code executed
in float64").

Shapes follow the reference: a multichannel spectrum is ``(M, F, T)``, a mask
``(F, T)``, a covariance ``(F, M, M)``, weights ``(F, M)`` here (the reference
carries them as ``(M, 1)`` columns per bin).
"""
from __future__ import annotations

import dataclasses
from typing import Callable, Optional

import numpy as np
import scipy.signal

__all__ = [
    "PathConfig", "PRESETS", "hann_periodic", "stft_scipy", "istft_scipy", "stft_np", "istft_np",
    "n_frames", "steering_vector", "all_steering_vectors", "ibm_noise_mask", "ibm_target_label",
    "geometric_phase_mask", "masked_covariance_loop", "masked_covariance_vec", "mvdr_weights",
    "beamform", "post_filter", "peak_normalise", "oracle_mask_mvdr", "geometric_mask_mvdr",
    "learned_mask_mvdr_chunk", "chunked_enhance", "batch_mvdr_vec", "logmag_ipd", "physics_features",
    "sir_sdr_unit_output", "osinr_osir", "far_field_delays", "fractional_delay", "mix_far_field",
    "streaming_mvdr", "hybrid_hard_null", "chunked_enhance_clipped",
]


# --------------------------------------------------------------------------------------
# configuration: one record of every constant in which the nine reference call sites differ
# (SURVEY.md 8-A2 "variant table")
# --------------------------------------------------------------------------------------
@dataclasses.dataclass(frozen=True)
class PathConfig:
    fs: float = 16000.0
    n_fft: int = 512
    hop: int = 128
    mic_dist: float = 0.01
    c: float = 343.0
    angle_deg: float = 90.0
    sigma: float = 1.0            # diagonal loading
    hp_hz: Optional[float] = 100.0  # bins with f < hp_hz: 'zero' or 'mic0'; None = no high-pass at all
    hp_mode: str = "zero"
    sqrt_eps: float = 0.0         # added to the noise weight inside the sqrt (tf_lite variant)
    norm_eps: float = 1e-6        # added to sum of weights
    w_eps: float = 1e-10          # added to d^H u
    post: str = "one_minus_noise"  # one_minus_noise | floor | mask | none
    post_floor: float = 0.05
    peak_eps: Optional[float] = 0.0  # None: no peak normalisation; else x / (max|x| + peak_eps)


PRESETS = {
    # rt_av_zoom/core/oracle_debug.py:11-24,42-94 at the BASELINE STFT shape (nb cell6:31-32)
    "baseline_oracle": PathConfig(),
    # rt_av_zoom/core/oracle_debug.py as written (N_HOP=256 passed as noverlap)
    "oracle_debug": PathConfig(hop=256),
    # rt_av_zoom/core/masked_mvdr.py:9-18,76-128
    "masked_mvdr": PathConfig(hop=256, sigma=1e-7, post="none", peak_eps=1e-6),
    # rt_av_zoom/core/full_audio_generating_pipeline/inference.py:18-26,88-118 with config.json
    "full_audio": PathConfig(n_fft=1024, hop=512, mic_dist=0.04, sigma=1e-5, post="floor", peak_eps=None),
    # rt_av_zoom/core/tf_lite_version/inference.py:85-179,349,375
    "tf_lite": PathConfig(n_fft=1024, hop=512, mic_dist=0.04, sigma=1e-5, hp_hz=None, sqrt_eps=1e-10,
                          post="floor", peak_eps=1e-9),
}


# --------------------------------------------------------------------------------------
# STFT / iSTFT  (rows 1 and 9 of SURVEY.md 8-A)
# --------------------------------------------------------------------------------------
def hann_periodic(n: int) -> np.ndarray:
    """scipy.signal.get_window('hann', n) (fftbins=True): w[j] = 0.5 - 0.5 cos(2 pi j / n)."""
    return 0.5 - 0.5 * np.cos(2.0 * np.pi * np.arange(n) / n)


def n_frames(length: int, n_fft: int, hop: int) -> int:
    """Frame count of scipy.signal.stft(boundary='zeros', padded=True) for a length-`length` input."""
    ext = length + 2 * (n_fft // 2)
    nadd = (-(ext - n_fft) % hop) % n_fft
    return (ext + nadd - n_fft) // hop + 1


def stft_scipy(x, n_fft: int, hop: int, fs: float = 16000.0) -> np.ndarray:
    """The call every reference site makes: oracle_debug.py:42-44, masked_mvdr.py:76,
    full_audio_generating_pipeline/inference.py:90 (noverlap = n_fft - hop)."""
    x = np.asarray(x, dtype=np.float64)
    return scipy.signal.stft(x, fs=fs, nperseg=n_fft, noverlap=n_fft - hop)[2]


def istft_scipy(Z, n_fft: int, hop: int, fs: float = 16000.0) -> np.ndarray:
    """oracle_debug.py:93, masked_mvdr.py:127, full_audio.../inference.py:117."""
    Z = np.asarray(Z, dtype=np.complex128)
    return scipy.signal.istft(Z, fs=fs, nperseg=n_fft, noverlap=n_fft - hop)[1]


def stft_np(x, n_fft: int, hop: int) -> np.ndarray:
    """Explicit restatement of what scipy's legacy `_spectral_helper` does for the call above
    (scipy/signal/_spectral_py.py: zero extension by n/2, tail zero-pad to a whole number of
    hops, periodic Hann, rfft, scale 1/sum(w)).  x: (..., L) -> (..., F, T)."""
    x = np.asarray(x, dtype=np.float64)
    L = x.shape[-1]
    half = n_fft // 2
    T = n_frames(L, n_fft, hop)
    ext_len = (T - 1) * hop + n_fft
    xe = np.zeros(x.shape[:-1] + (ext_len,))
    xe[..., half:half + L] = x
    w = hann_periodic(n_fft)
    idx = np.arange(T)[:, None] * hop + np.arange(n_fft)[None, :]
    frames = xe[..., idx] * w                      # (..., T, n_fft)
    Z = np.fft.rfft(frames, axis=-1) / w.sum()     # (..., T, F)
    return np.swapaxes(Z, -1, -2)


def istft_np(Z, n_fft: int, hop: int) -> np.ndarray:
    """Explicit restatement of scipy.signal.istft for the reference's call: irfft per frame
    (imaginary parts of the DC and Nyquist bins are ignored by the c2r transform), times sum(w),
    times w, overlap-add, divide by sum of w^2 where it exceeds 1e-10, drop n/2 at both ends.
    Z: (..., F, T) -> (..., (T-1)*hop)."""
    Z = np.asarray(Z, dtype=np.complex128)
    T = Z.shape[-1]
    w = hann_periodic(n_fft)
    xs = np.fft.irfft(np.swapaxes(Z, -1, -2), n=n_fft, axis=-1) * w.sum()   # (..., T, n_fft)
    out_len = n_fft + (T - 1) * hop
    x = np.zeros(Z.shape[:-2] + (out_len,))
    norm = np.zeros(out_len)
    for t in range(T):
        x[..., t * hop:t * hop + n_fft] += xs[..., t, :] * w
        norm[t * hop:t * hop + n_fft] += w * w
    half = n_fft // 2
    x = x[..., half:out_len - half]
    norm = norm[half:out_len - half]
    return x / np.where(norm > 1e-10, norm, 1.0)


# --------------------------------------------------------------------------------------
# steering vectors (row 5)
# --------------------------------------------------------------------------------------
def steering_vector(angle_deg: float, f: float, d: float, c: float) -> np.ndarray:
    """masked_mvdr.py:22-35 -> (2,) complex128: [exp(-i w tau1), exp(-i w tau2)],
    tau1 = (d/2) cos(theta)/c, tau2 = (d/2) cos(theta - pi)/c."""
    theta = np.deg2rad(angle_deg)
    tau1 = (d / 2) * np.cos(0.0) * np.cos(theta - 0) / c
    tau2 = (d / 2) * np.cos(0.0) * np.cos(theta - np.pi) / c
    omega = 2 * np.pi * f
    return np.array([np.exp(-1j * omega * tau1), np.exp(-1j * omega * tau2)], dtype=complex)


def all_steering_vectors(f_bins, angle_deg: float, d: float, c: float) -> np.ndarray:
    """tf_lite_version/inference.py:53-81 -> (F, 2) complex128."""
    f_bins = np.asarray(f_bins, dtype=np.float64)
    theta = np.deg2rad(angle_deg)
    tau1 = (d / 2) * np.cos(theta) / c
    tau2 = (d / 2) * np.cos(theta - np.pi) / c
    omega = 2 * np.pi * f_bins
    return np.stack([np.exp(-1j * omega * tau1), np.exp(-1j * omega * tau2)], axis=1)


# --------------------------------------------------------------------------------------
# masks (rows 3, 3b)
# --------------------------------------------------------------------------------------
def ibm_noise_mask(S_tgt, S_int) -> np.ndarray:
    """oracle_debug.py:49-53: 1.0 where |S_int| > |S_tgt| (strict), else 0.0."""
    return np.where(np.abs(np.asarray(S_int)) > np.abs(np.asarray(S_tgt)), 1.0, 0.0)


def ibm_target_label(S_tgt, S_int) -> np.ndarray:
    """model_training.py:90: training label |S_t| > |S_i| as float32 (ties -> 0 in both polarities)."""
    return (np.abs(np.asarray(S_tgt)) > np.abs(np.asarray(S_int))).astype(np.float32)


def geometric_phase_mask(Y) -> np.ndarray:
    """masked_mvdr.py:37-46: 1.0 where |angle(Y0) - angle(Y1)| > 0 else 0.01."""
    Y = np.asarray(Y)
    phase_diff = np.angle(Y[0]) - np.angle(Y[1])
    return np.where(np.abs(phase_diff) > 0.0, 1.0, 0.01)


# --------------------------------------------------------------------------------------
# covariance / weights / beamform (rows 4, 6, 7, 8)
# --------------------------------------------------------------------------------------
def masked_covariance_loop(Y, noise_w, sqrt_eps: float = 0.0, norm_eps: float = 1e-6) -> np.ndarray:
    """oracle_debug.py:56-64 / masked_mvdr.py:92-102, one bin at a time exactly as written:
    R[f] = (Y sqrt(m)) (Y sqrt(m))^H / (sum(m) + norm_eps).  Y (M,F,T), noise_w (F,T) -> (F,M,M)."""
    Y = np.asarray(Y, dtype=np.complex128)
    noise_w = np.asarray(noise_w, dtype=np.float64)
    M, F, _ = Y.shape
    R = np.zeros((F, M, M), dtype=complex)
    for k in range(F):
        m_f = noise_w[k, :]
        Yw = Y[:, k, :] * np.sqrt(m_f + sqrt_eps)
        R[k] = (Yw @ Yw.conj().T) / (np.sum(m_f) + norm_eps)
    return R


def masked_covariance_vec(Y, noise_w, sqrt_eps: float = 0.0, norm_eps: float = 1e-6) -> np.ndarray:
    """tf_lite_version/inference.py:97-127, all bins at once (einsum 'fmt,fnt->fmn')."""
    Y = np.asarray(Y, dtype=np.complex128)
    noise_w = np.asarray(noise_w)          # dtype kept: a float32 mask makes sqrt() and sum() float32 in the reference
    Yp = np.transpose(Y, (1, 0, 2))
    Yw = Yp * np.sqrt(noise_w[:, None, :] + sqrt_eps)
    R = np.einsum("fmt,fnt->fmn", Yw, Yw.conj())
    return R / (np.sum(noise_w, axis=1)[:, None, None] + norm_eps)


def mvdr_weights(R, d, sigma: float, w_eps: float = 1e-10) -> np.ndarray:
    """oracle_debug.py:70-79: per bin u = solve(R + sigma I, d); w = u / (d^H u + w_eps);
    LinAlgError (exactly singular) -> w = [1, 0].  R (F,M,M), d (F,M) -> (F,M)."""
    R = np.asarray(R, dtype=np.complex128)
    d = np.asarray(d, dtype=np.complex128)
    F, M, _ = R.shape
    w = np.zeros((F, M), dtype=complex)
    for k in range(F):
        Rl = R[k] + sigma * np.eye(M)
        dk = d[k].reshape(M, 1)
        try:
            u = np.linalg.solve(Rl, dk)
            u = u / (dk.conj().T @ u + w_eps)
        except np.linalg.LinAlgError:
            u = np.zeros((M, 1), dtype=complex)
            u[0, 0] = 1.0
        w[k] = u[:, 0]
    return w


def beamform(w, Y) -> np.ndarray:
    """oracle_debug.py:80: S[f,t] = sum_m conj(w[f,m]) Y[m,f,t]."""
    return np.einsum("fm,mft->ft", np.conj(np.asarray(w)), np.asarray(Y))


def post_filter(S, mask, cfg: PathConfig) -> np.ndarray:
    """oracle_debug.py:84-90 (mask = noise IBM, gain = 1 - mask);
    full_audio.../inference.py:116 (mask = target probability, gain = max(mask, floor));
    Final_pipeline/src/inference.py:219 (gain = mask)."""
    if cfg.post == "one_minus_noise":
        return S * (1.0 - mask)
    if cfg.post == "floor":
        return S * np.maximum(mask, cfg.post_floor)
    if cfg.post == "mask":
        return S * mask
    if cfg.post == "none":
        return S
    raise ValueError(cfg.post)


def peak_normalise(x, peak_eps: Optional[float]) -> np.ndarray:
    """oracle_debug.py:94 (eps 0), masked_mvdr.py:128 (1e-6), tf_lite.../inference.py:375 (1e-9)."""
    if peak_eps is None:
        return x
    return x / (np.max(np.abs(x)) + peak_eps)


def _freqs(cfg: PathConfig) -> np.ndarray:
    return np.fft.rfftfreq(cfg.n_fft, 1.0 / cfg.fs)


def _mvdr_from_noise_weight(Y, noise_w, cfg: PathConfig) -> np.ndarray:
    """Shared middle of every loop-form site: covariance -> weights -> beamform with the
    high-pass skip (oracle_debug.py:56-80; full_audio.../inference.py:102-114)."""
    f = _freqs(cfg)
    F = Y.shape[1]
    R = masked_covariance_loop(Y, noise_w, cfg.sqrt_eps, cfg.norm_eps)
    d = np.stack([steering_vector(cfg.angle_deg, f[k], cfg.mic_dist, cfg.c) for k in range(F)])
    w = mvdr_weights(R, d, cfg.sigma, cfg.w_eps)
    S = beamform(w, Y)
    if cfg.hp_hz is not None:
        low = f < cfg.hp_hz
        if cfg.hp_mode == "zero":
            S[low, :] = 0.0
        elif cfg.hp_mode == "mic0":
            S[low, :] = Y[0][low, :]
        else:
            raise ValueError(cfg.hp_mode)
    return S


def oracle_mask_mvdr(mix, tgt, itf, cfg: PathConfig = PRESETS["baseline_oracle"], return_parts: bool = False):
    """oracle_debug.py:42-94 end to end.  mix (2,L), tgt (L,), itf (L,) -> waveform ((T-1)*hop,)."""
    Y = stft_scipy(mix, cfg.n_fft, cfg.hop, cfg.fs)
    S_t = stft_scipy(tgt, cfg.n_fft, cfg.hop, cfg.fs)
    S_i = stft_scipy(itf, cfg.n_fft, cfg.hop, cfg.fs)
    mask_noise = ibm_noise_mask(S_t, S_i)
    S = _mvdr_from_noise_weight(Y, mask_noise, cfg)
    S_final = post_filter(S, mask_noise, cfg)
    x = istft_scipy(S_final, cfg.n_fft, cfg.hop, cfg.fs)
    out = peak_normalise(x, cfg.peak_eps)
    if return_parts:
        return out, {"Y": Y, "mask_noise": mask_noise, "S": S, "x_raw": x}
    return out


def geometric_mask_mvdr(mix, cfg: PathConfig = PRESETS["masked_mvdr"]):
    """masked_mvdr.py:76-128: geometric phase mask -> covariance -> MVDR -> iSTFT -> peak norm."""
    Y = stft_scipy(mix, cfg.n_fft, cfg.hop, cfg.fs)
    mask_noise = geometric_phase_mask(Y)
    S = _mvdr_from_noise_weight(Y, mask_noise, cfg)
    x = istft_scipy(S, cfg.n_fft, cfg.hop, cfg.fs)
    return peak_normalise(x, cfg.peak_eps)


def batch_mvdr_vec(Y, mask, f_bins, d_vectors, sigma: float) -> np.ndarray:
    """tf_lite_version/inference.py:85-179 restated: noise weight 1 - mask, sqrt eps 1e-10,
    norm eps 1e-6, broadcast solve, w eps 1e-10, NO high-pass.  d_vectors (F,2,1) -> (F,T)."""
    Y = np.asarray(Y, dtype=np.complex128)
    noise_w = 1.0 - np.asarray(mask)       # `mask_noise = (1.0 - mask)` in the mask's own dtype (:100)
    R = masked_covariance_vec(Y, noise_w, sqrt_eps=1e-10, norm_eps=1e-6)
    R = R + sigma * np.eye(2)[None]
    dv = np.asarray(d_vectors, dtype=np.complex128)
    try:
        u = np.linalg.solve(R, dv)
    except np.linalg.LinAlgError:
        u = np.zeros_like(dv)
        u[:, 0, :] = 1.0
    denom = np.matmul(np.transpose(dv.conj(), (0, 2, 1)), u) + 1e-10
    w = u / denom
    return np.matmul(np.transpose(w.conj(), (0, 2, 1)), np.transpose(Y, (1, 0, 2)))[:, 0, :]


def hybrid_hard_null(Y, mask, f_bins, mic_dist: float = 0.08, c: float = 343.0, angle_deg: float = 90.0) -> np.ndarray:
    """Final_pipeline/src/inference.py:28-98 (`hybrid_hard_null_bf`) restated: per bin, interference covariance
    R = (Y m)(Y)^H / (sum m + 1e-6) with m = 1 - mask; principal eigenvector (eigh) phase-normalised to mic 0; target
    steering vector normalised to mic 0; constraint matrix C = [v_tgt, v_int]; 2-norm condition number > 10 ->
    delay-and-sum w = v_tgt / 2, else solve C^H w = [1, 0]; bins below 200 Hz pass mic 0.  """
    Y = np.asarray(Y, dtype=np.complex128)
    m_int = 1.0 - np.asarray(mask)         # `mask_int = 1.0 - mask` in the mask's own dtype (:43)
    F, T = Y.shape[1], Y.shape[2]
    S = np.zeros((F, T), dtype=complex)
    e1 = np.array([[1], [0]], dtype=np.complex64)
    for i in range(F):
        f_hz = f_bins[i]
        if f_hz < 200:
            S[i, :] = Y[0, i, :]
            continue
        Yv = Y[:, i, :]
        mv = m_int[i, :]
        R = (Yv * mv) @ (Yv.conj().T) / (np.sum(mv) + 1e-6)
        _, vecs = np.linalg.eigh(R)
        v_int = vecs[:, -1].reshape(2, 1)
        v_int = v_int / (v_int[0] / (np.abs(v_int[0]) + 1e-10))
        theta = np.deg2rad(angle_deg)
        tau1 = (mic_dist / 2) * np.cos(theta) / c
        tau2 = (mic_dist / 2) * np.cos(theta - np.pi) / c
        om = 2 * np.pi * f_hz
        v_tgt = np.array([[np.exp(-1j * om * tau1)], [np.exp(-1j * om * tau2)]])
        v_tgt = v_tgt / (v_tgt[0] + 1e-10)
        Cm = np.column_stack((v_tgt, v_int))
        if np.linalg.cond(Cm) > 10:
            w = v_tgt / 2
        else:
            try:
                w = np.linalg.solve(Cm.conj().T, e1)
            except np.linalg.LinAlgError:
                w = v_tgt / 2
        S[i, :] = (w.conj().T @ Yv).squeeze()
    return S


# --------------------------------------------------------------------------------------
# features (row 2)
# --------------------------------------------------------------------------------------
def logmag_ipd(Y) -> np.ndarray:
    """full_audio.../inference.py:91-94: stack[ln(|Y0| + 1e-7), angle(Y0) - angle(Y1)] -> (2,F,T) f32."""
    Y = np.asarray(Y)
    mag0 = np.abs(Y[0])
    ipd = np.angle(Y[0]) - np.angle(Y[1])
    return np.stack([np.log(mag0 + 1e-7), ipd], axis=0).astype(np.float32)


def physics_features(Y, n_fft: int) -> np.ndarray:
    """Final_pipeline/src/inference.py:202-204,117-128: (F,T,4) f32 NHWC
    [logmag, sin ipd, cos ipd, linspace(0,1,F) broadcast]."""
    Y = np.asarray(Y)
    F, T = Y.shape[1], Y.shape[2]
    log_mag = np.log(np.abs(Y[0]) + 1e-7)
    ipd = np.angle(Y[0]) - np.angle(Y[1])
    fmap = np.tile(np.linspace(0, 1, F).reshape(F, 1), (1, T))
    return np.stack([log_mag, np.sin(ipd), np.cos(ipd), fmap], axis=-1).astype(np.float32)


# --------------------------------------------------------------------------------------
# learned-mask chunk path (rows 2-9, 9b)
# --------------------------------------------------------------------------------------
def learned_mask_mvdr_chunk(y_chunk, mask_fn: Callable[[np.ndarray], np.ndarray],
                            cfg: PathConfig = PRESETS["full_audio"]) -> np.ndarray:
    """full_audio.../inference.py:88-118 (`process_chunk`).  y_chunk (N,2); mask_fn maps the
    (2,F,T) float32 feature stack to a (F,T) target-probability mask."""
    y_chunk = np.asarray(y_chunk, dtype=np.float64)
    Y = stft_scipy(y_chunk.T, cfg.n_fft, cfg.hop, cfg.fs)
    mask = np.asarray(mask_fn(logmag_ipd(Y)), dtype=np.float64)
    S = _mvdr_from_noise_weight(Y, 1.0 - mask, cfg)
    S_final = post_filter(S, mask, cfg)
    return istft_scipy(S_final, cfg.n_fft, cfg.hop, cfg.fs)


def chunked_enhance(y_full, mask_fn, cfg: PathConfig = PRESETS["full_audio"], win: int = 32000) -> np.ndarray:
    """full_audio.../inference.py:127-156 (`main_deploy` body): windows of `win` at stride win/2,
    ceil(L/stride) windows, zero-padded tail, count-averaged overlap-add, trimmed to L."""
    y_full = np.asarray(y_full, dtype=np.float64)
    L = y_full.shape[0]
    stride = win // 2
    out_buf = np.zeros(L + win)
    cnt_buf = np.zeros(L + win)
    n_chunks = int(np.ceil(L / stride))
    for i in range(n_chunks):
        s = i * stride
        chunk = y_full[s:s + win]
        if chunk.shape[0] < win:
            chunk = np.pad(chunk, ((0, win - chunk.shape[0]), (0, 0)))
        o = learned_mask_mvdr_chunk(chunk, mask_fn, cfg)
        n = min(o.shape[0], win)
        out_buf[s:s + n] += o[:n]
        cnt_buf[s:s + n] += 1.0
    cnt_buf[cnt_buf == 0] = 1.0
    return out_buf[:L] / cnt_buf[:L]


def chunked_enhance_clipped(y_full, mask_fn, beamformer: str = "batch_mvdr", win: int = 32000, fs: float = 16000.0,
                            n_fft: int = 1024, hop: int = 512, mic_dist: float = 0.04, c: float = 343.0,
                            sigma: float = 1e-5, peak_eps: float = 1e-9) -> np.ndarray:
    """The TFLite-era chunk drivers restated: `process_audio_file` (tf_lite_version/inference.py:245-391,
    beamformer='batch_mvdr', post-filter max(mask, 0.05), d = 0.04) and `enhance_audio`
    (Final_pipeline/src/inference.py:144-238, beamformer='hybrid_null', post-filter mask, MIC_DIST = 0.08).
    Buffers of len(y): a window adds all its iSTFT samples up to the end of the buffer
    (w_len = min(len(chunk_out), len(out_buf[start:]))), count-averaged, then / (max|x| + 1e-9).
    mask_fn(log_mag (F,T), ipd (F,T)) -> (F,T) target-probability mask (the TFLite interpreter's role)."""
    y_full = np.asarray(y_full)
    L = y_full.shape[0]
    stride = win // 2
    out_buf = np.zeros(L)
    norm_buf = np.zeros(L)
    for i in range(int(np.ceil(L / stride))):
        s = i * stride
        chunk = y_full[s:s + win]
        if chunk.shape[0] < win:
            chunk = np.pad(chunk, ((0, win - chunk.shape[0]), (0, 0)))
        f_bins, _, Y = scipy.signal.stft(chunk.T, fs=fs, nperseg=n_fft, noverlap=n_fft - hop)
        mask = np.asarray(mask_fn(np.log(np.abs(Y[0]) + 1e-7), np.angle(Y[0]) - np.angle(Y[1])))
        if beamformer == "batch_mvdr":
            d_vecs = all_steering_vectors(f_bins, 90.0, mic_dist, c)[:, :, None]
            S_final = batch_mvdr_vec(Y, mask, f_bins, d_vecs, sigma) * np.maximum(mask, 0.05)
        elif beamformer == "hybrid_null":
            S_final = hybrid_hard_null(Y, mask, f_bins, mic_dist, c) * mask
        else:
            raise ValueError(beamformer)
        _, chunk_out = scipy.signal.istft(S_final, fs=fs, nperseg=n_fft, noverlap=n_fft - hop)
        w_len = min(len(chunk_out), L - s)
        out_buf[s:s + w_len] += chunk_out[:w_len]
        norm_buf[s:s + w_len] += 1.0
    final = out_buf / np.maximum(norm_buf, 1.0)
    return final / (np.max(np.abs(final)) + peak_eps)


# --------------------------------------------------------------------------------------
# scores (row 11)
# --------------------------------------------------------------------------------------
def sir_sdr_unit_output(output, target, interf):
    """scripts/run_metrics.py:6-36 (`calculate_metrics_manual`) -> (sdr, sir) in dB."""
    eps = 1e-10
    o = np.asarray(output, dtype=np.float64)
    t = np.asarray(target, dtype=np.float64)
    i = np.asarray(interf, dtype=np.float64)
    o = o / (np.linalg.norm(o) + eps)
    t = t / (np.linalg.norm(t) + eps)
    i = i / (np.linalg.norm(i) + eps)
    e_t = np.dot(o, t) * t
    e_i = np.dot(o, i) * i
    e_a = o - e_t - e_i
    p_t = np.sum(e_t ** 2)
    p_i = np.sum(e_i ** 2) + 1e-10
    p_n = np.sum(e_a ** 2) + 1e-10
    return 10 * np.log10(p_t / (p_i + p_n)), 10 * np.log10(p_t / p_i)


def osinr_osir(output, target, interferer):
    """Final_pipeline/src/metrics.py:102-123 (`calculate_osnr_osir`) -> (OSINR, OSIR) in dB."""
    eps = 1e-10
    o = np.asarray(output, dtype=np.float64)
    t = np.asarray(target, dtype=np.float64)
    i = np.asarray(interferer, dtype=np.float64)
    t = t / (np.linalg.norm(t) + eps)
    i = i / (np.linalg.norm(i) + eps)
    e_t = np.dot(o, t) * t
    e_i = np.dot(o, i) * i
    e_n = o - e_t - e_i
    P_t, P_i, P_n = np.sum(e_t ** 2), np.sum(e_i ** 2), np.sum(e_n ** 2)
    return 10 * np.log10(P_t / (P_i + P_n + eps)), 10 * np.log10(P_t / (P_i + eps))


# --------------------------------------------------------------------------------------
# far-field mixer (world_building; input recipe for the benchmark)
# --------------------------------------------------------------------------------------
def far_field_delays(az_deg: float, d: float, c: float):
    """tf_lite_version/world_building.py:40-44."""
    th = np.deg2rad(az_deg)
    return (d / 2) * np.cos(th - 0) / c, (d / 2) * np.cos(th - np.pi) / c


def fractional_delay(y, delay_sec: float, fs: float) -> np.ndarray:
    """tf_lite_version/world_building.py:46-52: whole-signal rFFT phase ramp."""
    y = np.asarray(y, dtype=np.float64)
    n = len(y)
    spec = np.fft.rfft(y)
    fr = np.fft.rfftfreq(n, 1.0 / fs)
    return np.fft.irfft(spec * np.exp(-1j * 2 * np.pi * fr * delay_sec), n=n)


def mix_far_field(sources, angles_deg, d: float = 0.04, c: float = 343.0, fs: float = 16000.0):
    """tf_lite_version/world_building.py:61-93 (`mix_and_save` without the file I/O): source 0 is
    the target, the rest interferers; references are the mic-1 images; all three are divided by
    max|mix| + 1e-9.  -> mix (2,L), tgt_ref (L,), int_ref (L,) float64."""
    L = max(len(s) for s in sources)
    m1 = np.zeros(L)
    m2 = np.zeros(L)
    tgt = np.zeros(L)
    itf = np.zeros(L)
    for idx, (s, a) in enumerate(zip(sources, angles_deg)):
        s = np.pad(np.asarray(s, dtype=np.float64), (0, L - len(s)))
        t1, t2 = far_field_delays(a, d, c)
        s1 = fractional_delay(s, t1, fs)
        s2 = fractional_delay(s, t2, fs)
        m1 += s1
        m2 += s2
        if idx == 0:
            tgt += s1
        else:
            itf += s1
    mix = np.stack([m1, m2], axis=0)
    norm = np.max(np.abs(mix)) + 1e-9
    return mix / norm, tgt / norm, itf / norm


# --------------------------------------------------------------------------------------
# streaming recursion (row 10)  -- NOT IN THE REFERENCE: parity unpinned
# --------------------------------------------------------------------------------------
def streaming_mvdr(mix, noise_w_fn, cfg: PathConfig, lam: float = 0.95):
    """Recursive exponentially-smoothed variant defined by this project:
        R_t = lam R_{t-1} + (1-lam) m_t y_t y_t^H ,  n_t = lam n_{t-1} + (1-lam) m_t
        w_t = mvdr(R_t / (n_t + norm_eps) + sigma I) ,  S_t = w_t^H y_t  (bins below hp_hz -> 0)
    followed by the same iSTFT.  `noise_w_fn(Y_t (2,F)) -> (F,)` supplies the per-frame noise
    weight.  Parity unpinned: there is no reference implementation of this recursion."""
    Y = stft_scipy(mix, cfg.n_fft, cfg.hop, cfg.fs)
    M, F, T = Y.shape
    f = _freqs(cfg)
    d = np.stack([steering_vector(cfg.angle_deg, f[k], cfg.mic_dist, cfg.c) for k in range(F)])
    R = np.zeros((F, M, M), dtype=complex)
    n = np.zeros(F)
    S = np.zeros((F, T), dtype=complex)
    for t in range(T):
        y = Y[:, :, t]                       # (M,F)
        m = np.asarray(noise_w_fn(y), dtype=np.float64)
        R = lam * R + (1 - lam) * m[:, None, None] * np.einsum("mf,nf->fmn", y, y.conj())
        n = lam * n + (1 - lam) * m
        w = mvdr_weights(R / (n[:, None, None] + cfg.norm_eps), d, cfg.sigma, cfg.w_eps)
        S[:, t] = np.einsum("fm,mf->f", w.conj(), y)
    if cfg.hp_hz is not None:
        S[f < cfg.hp_hz, :] = 0.0
    return istft_scipy(S, cfg.n_fft, cfg.hop, cfg.fs)

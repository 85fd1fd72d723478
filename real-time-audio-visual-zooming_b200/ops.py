"""Tensor-level wrappers over the C ABI (include/avzoom.h).

Every function takes CUDA `torch.Tensor`s (or numpy arrays, which are copied to the current CUDA device
and whose results are copied back - the reference's calling convention) and launches hand-written
sm_100a kernels on torch's current stream.  PyTorch is used only for device memory and streams.
Shapes follow the reference: spectra `(..., M, F, T)`, masks `(..., F, T)`; any leading dimensions are
the batch.  There is no CPU path.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import numpy as np
import torch

from . import _lib
from .config import MvdrConfig, PRESETS

__all__ = [
    "num_frames", "stft", "istft", "ibm", "ibm_target_label", "geometric_mask", "masked_covariance",
    "steering_vectors", "mvdr_weights", "hybrid_null_weights", "beamform", "logmag_ipd", "physics_features", "sir_scores",
    "ibm_covariance", "wave_masked_covariance", "mvdr_apply", "peak_normalise", "unpack_ibm",
    "oracle_mask_mvdr", "learned_mask_mvdr", "covariance_to_matrix", "ibm_exact_bits", "irm", "wave_features", "alloc_kept_spectrum",
    "far_field_mix", "fractional_delay", "pcm16_to_float", "float_to_pcm16",
]


# ------------------------------------------------------------------------------------------ helpers
def _stream() -> C.c_void_p:
    if not torch.cuda.is_available():
        raise _lib.AvzError("avzoom needs a CUDA device (sm_100a); there is no CPU fallback")
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t: Optional[torch.Tensor]) -> C.c_void_p:
    return C.c_void_p(0 if t is None else t.data_ptr())


class _Io:
    """Remembers whether the caller handed numpy (-> give numpy back) and moves data to the GPU."""

    def __init__(self):
        self.numpy = False

    def take(self, x, dtype) -> torch.Tensor:
        if isinstance(x, np.ndarray):
            self.numpy = True
            x = torch.from_numpy(np.ascontiguousarray(x))
        if not isinstance(x, torch.Tensor):
            x = torch.as_tensor(x)
        if not x.is_cuda:
            if not torch.cuda.is_available():
                raise _lib.AvzError("avzoom needs a CUDA device (sm_100a); there is no CPU fallback")
            x = x.cuda(non_blocking=True)
        if x.dtype != dtype:
            x = x.to(dtype)
        return x.contiguous()

    def give(self, t: torch.Tensor):
        return t.cpu().numpy() if self.numpy else t


_WIDE = (np.dtype(np.float64), np.dtype(np.complex128), torch.float64, torch.complex128)


def _wide(*xs) -> bool:
    """True if any argument is float64 / complex128: the call then runs on the float64 kernels (avz_*_f64), as numpy
    and scipy follow their input dtype (the reference's R, w, S are complex128: oracle_debug.py:57,67)."""
    for x in xs:
        dt = getattr(x, "dtype", None)
        if dt is not None and dt in _WIDE:
            return True
    return False


def num_frames(length: int, n_fft: int, hop: int) -> int:
    """Frame count of scipy.signal.stft(boundary='zeros', padded=True): ceil(L/hop) + 1 when hop | n_fft."""
    return int(_lib.load().avz_num_frames(int(length), int(n_fft), int(hop)))


def stft(x, n_fft: int = 512, hop: int = 128):
    """scipy.signal.stft(x, fs, nperseg=n_fft, noverlap=n_fft-hop)[2] (oracle_debug.py:42-44).
    x (..., L) float32 -> (..., F, T) complex64.  Pairs of rows along the second-to-last axis share one
    complex transform."""
    io = _Io()
    wide = _wide(x)
    x = io.take(x, torch.float64 if wide else torch.float32)
    lead = x.shape[:-1]
    L = x.shape[-1]
    if x.dim() == 1:
        B, Cn = 1, 1
    else:
        Cn = x.shape[-2]
        B = int(np.prod(x.shape[:-2])) if x.dim() > 2 else 1
    F, T = n_fft // 2 + 1, num_frames(L, n_fft, hop)
    Y = torch.empty((B, Cn, F, T), dtype=torch.complex128 if wide else torch.complex64, device=x.device)
    lib = _lib.load()
    if wide:    # float64 in -> complex128 out, like scipy
        _lib.check(lib.avz_stft_f64(_ptr(x), B, Cn, L, n_fft, hop, _ptr(Y), _stream()), "avz_stft_f64")
    else:
        _lib.check(lib.avz_stft_f32(_ptr(x), B, Cn, L, n_fft, hop, _ptr(Y), _stream()), "avz_stft_f32")
    return io.give(Y.reshape(*lead, F, T))


def istft(S, n_fft: int = 512, hop: int = 128, return_peak: bool = False):
    """scipy.signal.istft(S, fs, nperseg=n_fft, noverlap=n_fft-hop)[1] (oracle_debug.py:93).
    S (..., F, T) complex64 -> (..., (T-1)*hop) float32."""
    io = _Io()
    wide = _wide(S)
    S = io.take(S, torch.complex128 if wide else torch.complex64)
    F, T = S.shape[-2], S.shape[-1]
    if F != n_fft // 2 + 1:
        raise ValueError(f"spectrum has {F} bins, n_fft={n_fft} needs {n_fft // 2 + 1}")
    lead = S.shape[:-2]
    B = int(np.prod(lead)) if lead else 1
    if wide:
        lib = _lib.load()
        out = torch.empty((B, (T - 1) * hop), dtype=torch.float64, device=S.device)
        ws = torch.empty((int(lib.avz_istft_f64_ws_bytes(B, T, n_fft)),), dtype=torch.uint8, device=S.device)
        _lib.check(lib.avz_istft_f64(_ptr(S), B, T, n_fft, hop, _ptr(out), _ptr(ws), _stream()), "avz_istft_f64")
        out = out.reshape(*lead, (T - 1) * hop)
        if return_peak:
            return io.give(out), io.give(out.abs().amax(dim=-1))
        return io.give(out)
    out = torch.empty((B, (T - 1) * hop), dtype=torch.float32, device=S.device)
    peak = torch.zeros((B,), dtype=torch.float32, device=S.device) if return_peak else None
    lib = _lib.load()
    _lib.check(lib.avz_istft_f32(_ptr(S), B, T, n_fft, hop, _ptr(out), _ptr(peak), _stream()), "avz_istft_f32")
    out = out.reshape(*lead, (T - 1) * hop)
    if return_peak:
        return io.give(out), io.give(peak.reshape(lead))
    return io.give(out)


def peak_normalise(x: torch.Tensor, peak: torch.Tensor, peak_eps: float) -> torch.Tensor:
    """In place x[b] /= (peak[b] + peak_eps)  (oracle_debug.py:94).  float64 rows find their own maximum (`peak` unused)."""
    B, n = x.shape
    if x.dtype == torch.float64:
        _lib.check(_lib.load().avz_peak_normalise_f64(_ptr(x), B, n, float(peak_eps), _ptr(None), _stream()),
                   "avz_peak_normalise_f64")
        return x
    _lib.check(_lib.load().avz_peak_normalise_f32(_ptr(x), B, n, _ptr(peak), float(peak_eps), _stream()),
               "avz_peak_normalise_f32")
    return x


# ------------------------------------------------------------------------------------------ masks
def _mag_greater(a, b):
    io = _Io()
    a = io.take(a, torch.complex64)
    b = io.take(b, torch.complex64)
    if a.shape != b.shape:
        raise ValueError("spectra differ in shape")
    out = torch.empty(a.shape, dtype=torch.float32, device=a.device)
    _lib.check(_lib.load().avz_mag_greater_f32(_ptr(a), _ptr(b), a.numel(), _ptr(out), _stream()), "avz_mag_greater_f32")
    return io.give(out)


def ibm(S_tgt, S_int):
    """Noise-polarity Ideal Binary Mask: where(|S_int| > |S_tgt|, 1, 0)  (oracle_debug.py:49-53)."""
    return _mag_greater(S_int, S_tgt)


def ibm_target_label(S_tgt, S_int):
    """Training-label polarity: (|S_t| > |S_i|) as float  (model_training.py:90)."""
    return _mag_greater(S_tgt, S_int)


def irm(S_tgt, S_int):
    """Soft post-filter mask sqrt(Pt / (Pt + Pi + 1e-10))  (oracle_reverb.py:143-147)."""
    io = _Io()
    a = io.take(S_tgt, torch.complex64)
    b = io.take(S_int, torch.complex64)
    if a.shape != b.shape:
        raise ValueError("spectra differ in shape")
    out = torch.empty(a.shape, dtype=torch.float32, device=a.device)
    _lib.check(_lib.load().avz_irm_f32(_ptr(a), _ptr(b), a.numel(), _ptr(out), _stream()), "avz_irm_f32")
    return io.give(out)


def geometric_mask(Y):
    """compute_hard_geometric_mask (masked_mvdr.py:37-46).  Y (..., 2, F, T) -> (..., F, T) in {0.01, 1}."""
    io = _Io()
    wide = _wide(Y)
    Y = io.take(Y, torch.complex128 if wide else torch.complex64)
    F, T = Y.shape[-2], Y.shape[-1]
    lead = Y.shape[:-3]
    B = int(np.prod(lead)) if lead else 1
    out = torch.empty((B, F, T), dtype=torch.float64 if wide else torch.float32, device=Y.device)
    if wide:
        _lib.check(_lib.load().avz_geometric_mask_f64(_ptr(Y), B, F, T, _ptr(out), _stream()), "avz_geometric_mask_f64")
    else:
        _lib.check(_lib.load().avz_geometric_mask_f32(_ptr(Y), B, F, T, _ptr(out), _stream()), "avz_geometric_mask_f32")
    return io.give(out.reshape(*lead, F, T))


def unpack_ibm(bits: torch.Tensor, F: int) -> torch.Tensor:
    """ibm_bits [B, T, ceil(F/32)] int32 -> float mask [B, F, T]."""
    B, T, _ = bits.shape
    out = torch.empty((B, F, T), dtype=torch.float32, device=bits.device)
    _lib.check(_lib.load().avz_ibm_unpack_f32(_ptr(bits), B, F, T, _ptr(out), _stream()), "avz_ibm_unpack_f32")
    return out


# ------------------------------------------------------------------------------------------ covariance / weights
def covariance_to_matrix(Rp: torch.Tensor) -> torch.Tensor:
    """Packed (..., F, 4) = (R00, R11, Re R01, Im R01) -> Hermitian (..., F, 2, 2) complex64."""
    r01 = torch.complex(Rp[..., 2], Rp[..., 3])
    z = torch.zeros_like(Rp[..., 0])
    row0 = torch.stack([torch.complex(Rp[..., 0], z), r01], dim=-1)
    row1 = torch.stack([r01.conj(), torch.complex(Rp[..., 1], z)], dim=-1)
    return torch.stack([row0, row1], dim=-2)


def _pack_covariance(R: torch.Tensor, wide: bool = False) -> torch.Tensor:
    if R.shape[-1] == 4 and not R.is_complex():
        return R.to(torch.float64 if wide else torch.float32).contiguous()
    R = R.to(torch.complex128 if wide else torch.complex64)
    return torch.stack([R[..., 0, 0].real, R[..., 1, 1].real, R[..., 0, 1].real, R[..., 0, 1].imag], dim=-1).contiguous()


def masked_covariance(Y, noise_w, sqrt_eps: float = 0.0, norm_eps: float = 1e-6, packed: bool = False):
    """R[f] = sum_t (m + sqrt_eps) y y^H / (sum_t m + norm_eps)  (oracle_debug.py:56-64).
    Y (..., 2, F, T), noise_w (..., F, T) -> (..., F, 2, 2) complex64."""
    io = _Io()
    wide = _wide(Y)
    Y = io.take(Y, torch.complex128 if wide else torch.complex64)
    m = io.take(noise_w, torch.float64 if wide else torch.float32)
    F, T = Y.shape[-2], Y.shape[-1]
    lead = Y.shape[:-3]
    B = int(np.prod(lead)) if lead else 1
    if m.numel() != B * F * T:
        raise ValueError(f"noise weights {tuple(m.shape)} do not match the spectrum {tuple(Y.shape)}")
    Rp = torch.empty((B, F, 4), dtype=m.dtype, device=Y.device)
    ms = torch.empty((B, F), dtype=m.dtype, device=Y.device)
    if wide:
        _lib.check(_lib.load().avz_spec_mask_cov_f64(_ptr(Y), _ptr(m), B, F, T, float(sqrt_eps), float(norm_eps),
                                                     _ptr(Rp), _ptr(ms), _stream()), "avz_spec_mask_cov_f64")
    else:
        _lib.check(_lib.load().avz_spec_mask_cov_f32(_ptr(Y), _ptr(m), B, F, T, float(sqrt_eps), float(norm_eps), _ptr(Rp),
                                                     _ptr(ms), _stream()), "avz_spec_mask_cov_f32")
    Rp = Rp.reshape(*lead, F, 4)
    return io.give(Rp if packed else covariance_to_matrix(Rp))


_SV_CACHE = {}


def steering_vectors(cfg: MvdrConfig, device=None, f_bins=None, wide: bool = False) -> torch.Tensor:
    """get_steering_vector for every bin (masked_mvdr.py:22-35), float64 on the host, (F, 2) complex64 (`wide`: complex128)."""
    key = (cfg.angle_deg, cfg.mic_dist, cfg.c, cfg.fs, cfg.n_fft, str(device), None if f_bins is None else tuple(f_bins),
           wide)
    if key not in _SV_CACHE:
        f = cfg.freqs() if f_bins is None else np.asarray(f_bins, dtype=np.float64)
        th = np.deg2rad(cfg.angle_deg)
        tau1 = (cfg.mic_dist / 2) * np.cos(0.0) * np.cos(th - 0) / cfg.c
        tau2 = (cfg.mic_dist / 2) * np.cos(0.0) * np.cos(th - np.pi) / cfg.c
        om = 2 * np.pi * f
        d = np.stack([np.exp(-1j * om * tau1), np.exp(-1j * om * tau2)], axis=1).astype(np.complex128 if wide else np.complex64)
        _SV_CACHE[key] = torch.from_numpy(d).to(device if device is not None else "cuda")
    return _SV_CACHE[key]


def _take_cov_and_steering(io: "_Io", R, d, wide: bool):
    """R (..., F, 2, 2) complex or packed (..., F, 4); d (F, 2) or (F, 2, 1) -> packed R, d (F, 2), B, F, lead shape."""
    cdt, rdt = (torch.complex128, torch.float64) if wide else (torch.complex64, torch.float32)
    if isinstance(R, np.ndarray):
        r_dtype = cdt if np.iscomplexobj(R) else rdt
    else:
        r_dtype = cdt if R.is_complex() else rdt
    Rp = _pack_covariance(io.take(R, r_dtype), wide)
    d = io.take(d, cdt).reshape(-1, 2).contiguous()
    F = Rp.shape[-2]
    if d.shape[0] != F:
        raise ValueError(f"steering vectors cover {d.shape[0]} bins, the covariance {F}")
    lead = Rp.shape[:-2]
    B = int(np.prod(lead)) if lead else 1
    return Rp, d, B, F, lead


def mvdr_weights(R, d, cfg: MvdrConfig = PRESETS["baseline_oracle"]):
    """w = u / (d^H u + w_eps), u = solve(R + sigma I, d)  (oracle_debug.py:68-79), closed form in float64.
    R (..., F, 2, 2) complex or packed (..., F, 4); d (F, 2) or (F, 2, 1) -> w (..., F, 2) complex64.
    Bins below cfg.hp_hz get w = 0 ('zero') or [1, 0] ('mic0')."""
    io = _Io()
    wide = _wide(R)
    Rp, d, B, F, lead = _take_cov_and_steering(io, R, d, wide)
    w = torch.empty((B, F, 2), dtype=torch.complex128 if wide else torch.complex64, device=Rp.device)
    cc = cfg.to_c()
    if wide:    # sigma and w_eps travel as doubles (1e-7 is not a float32 number)
        _lib.check(_lib.load().avz_mvdr_weights_f64(_ptr(Rp), _ptr(d), B, F, float(cfg.sigma), float(cfg.w_eps),
                                                    int(cc.hp_bins), int(cc.hp_mode), _ptr(w), _stream()),
                   "avz_mvdr_weights_f64")
    else:
        _lib.check(_lib.load().avz_mvdr_weights_f32(_ptr(Rp), _ptr(d), B, F, C.byref(cc), _ptr(w), _stream()),
                   "avz_mvdr_weights_f32")
    return io.give(w.reshape(*lead, F, 2))


def hybrid_null_weights(R, d, bypass_bins: int, zero_cov_nan: bool = False, round_to_f32: bool = False):
    """Hybrid hard-null weights (Final_pipeline/src/inference.py:56-94) from the interference covariance.
    R (..., F, 2, 2) complex or packed (..., F, 4); d (F, 2) un-normalised steering vectors -> w (..., F, 2) complex64
    (complex128 for float64 / complex128 R); bins below `bypass_bins` pass mic 0.  `zero_cov_nan`: reproduce the
    reference's NaN for a bin whose principal eigenvector has no mic-0 component (default: delay-and-sum there)."""
    io = _Io()
    wide = _wide(R)
    Rp, d, B, F, lead = _take_cov_and_steering(io, R, d, wide)
    w = torch.empty((B, F, 2), dtype=torch.complex128 if (wide and not round_to_f32) else torch.complex64, device=Rp.device)
    lib = _lib.load()
    # round_to_f32: float64 arithmetic, weights rounded once to complex64 for the float32 pass B
    fn = (lib.avz_hybrid_null_weights_f64_w32 if round_to_f32 else lib.avz_hybrid_null_weights_f64) if wide \
        else lib.avz_hybrid_null_weights_f32
    _lib.check(fn(_ptr(Rp), _ptr(d), B, F, int(bypass_bins), int(bool(zero_cov_nan)), _ptr(w), _stream()),
               "avz_hybrid_null_weights")
    return io.give(w.reshape(*lead, F, 2))


def beamform(w, Y):
    """S[f,t] = w[f]^H Y[:,f,t]  (oracle_debug.py:80).  w (..., F, 2), Y (..., 2, F, T) -> (..., F, T)."""
    io = _Io()
    wide = _wide(w, Y)
    cdt = torch.complex128 if wide else torch.complex64
    w = io.take(w, cdt)
    Y = io.take(Y, cdt)
    F, T = Y.shape[-2], Y.shape[-1]
    lead = Y.shape[:-3]
    B = int(np.prod(lead)) if lead else 1
    if w.numel() != B * F * 2:
        raise ValueError(f"weights {tuple(w.shape)} do not match the spectrum {tuple(Y.shape)}")
    S = torch.empty((B, F, T), dtype=cdt, device=Y.device)
    if wide:
        _lib.check(_lib.load().avz_beamform_f64(_ptr(w), _ptr(Y), B, F, T, _ptr(S), _stream()), "avz_beamform_f64")
    else:
        _lib.check(_lib.load().avz_beamform_f32(_ptr(w), _ptr(Y), B, F, T, _ptr(S), _stream()), "avz_beamform_f32")
    return io.give(S.reshape(*lead, F, T))


# ------------------------------------------------------------------------------------------ features / scores
def _features(Y, mode: int):
    io = _Io()
    Y = io.take(Y, torch.complex64)
    F, T = Y.shape[-2], Y.shape[-1]
    lead = Y.shape[:-3]
    B = int(np.prod(lead)) if lead else 1
    if mode == _lib.FEAT_PHYSICS_NHWC:
        X = torch.empty((B, F, T, 4), dtype=torch.float32, device=Y.device)
        shape = (*lead, F, T, 4)
    else:
        X = torch.empty((B, 2, F, T), dtype=torch.float32, device=Y.device)
        shape = (*lead, 2, F, T)
    _lib.check(_lib.load().avz_features_f32(_ptr(Y), B, F, T, mode, _ptr(X), _stream()), "avz_features_f32")
    return io.give(X.reshape(shape))


def logmag_ipd(Y, wrapped: bool = False):
    """stack[ln(|Y0| + 1e-7), angle(Y0) - angle(Y1)]  (full_audio.../inference.py:91-94) -> (..., 2, F, T) f32."""
    return _features(Y, _lib.FEAT_LOGMAG_IPD_WRAPPED if wrapped else _lib.FEAT_LOGMAG_IPD)


def physics_features(Y):
    """[logmag, sin ipd, cos ipd, k/(F-1)] NHWC  (Final_pipeline/src/inference.py:117-128) -> (..., F, T, 4)."""
    return _features(Y, _lib.FEAT_PHYSICS_NHWC)


def wave_features(mix, n_fft: int = 1024, hop: int = 512, mode: str = "logmag_ipd"):
    """Features straight from the waveform, STFT fused (no spectrum written): full_audio.../inference.py:90-94.
    mix (..., 2, L) -> (..., 2, F, T) f32 ('logmag_ipd', 'logmag_ipd_wrapped') or (..., F, T, 4) ('physics')."""
    io = _Io()
    mix = io.take(mix, torch.float32)
    lead = mix.shape[:-2]
    B = int(np.prod(lead)) if lead else 1
    L = mix.shape[-1]
    F, T = n_fft // 2 + 1, num_frames(L, n_fft, hop)
    m = {"logmag_ipd": _lib.FEAT_LOGMAG_IPD, "logmag_ipd_wrapped": _lib.FEAT_LOGMAG_IPD_WRAPPED,
         "physics": _lib.FEAT_PHYSICS_NHWC}[mode]
    shape = (B, F, T, 4) if m == _lib.FEAT_PHYSICS_NHWC else (B, 2, F, T)
    X = torch.empty(shape, dtype=torch.float32, device=mix.device)
    _lib.check(_lib.load().avz_wave_features_f32(_ptr(mix), B, L, n_fft, hop, m, _ptr(X), _stream()),
               "avz_wave_features_f32")
    return io.give(X.reshape(*lead, *shape[1:]))


def sir_scores(est, tgt, itf):
    """Projection scores per utterance -> (..., 4) float32 = (OSINR, OSIR, SDR, SIR) dB:
    columns 0-1 follow Final_pipeline/src/metrics.py:102-123, columns 2-3 scripts/run_metrics.py:6-36."""
    io = _Io()
    est = io.take(est, torch.float32)
    tgt = io.take(tgt, torch.float32)
    itf = io.take(itf, torch.float32)
    lead = est.shape[:-1]
    B = int(np.prod(lead)) if lead else 1
    if tgt.shape != itf.shape or tgt.shape[:-1] != lead:
        raise ValueError(f"references {tuple(tgt.shape)} / {tuple(itf.shape)} do not match the estimate {tuple(est.shape)}")
    sc = torch.empty((B, 4), dtype=torch.float32, device=est.device)
    _lib.check(_lib.load().avz_sir_f32(_ptr(est), _ptr(tgt), _ptr(itf), B, est.shape[-1], tgt.shape[-1], _ptr(sc),
                                       _stream()), "avz_sir_f32")
    return io.give(sc.reshape(*lead, 4))


# ------------------------------------------------------------------------------------------ mixer / wire format
def far_field_mix(sources, delays_s, fs: float = 16000.0, peak_eps: Optional[float] = 1e-9, multi_pass: bool = False):
    """Far-field 2-mic mixtures on the device (tf_lite_version/world_building.py:61-93 without the file I/O).

    sources (B,S,L) or (S,L) float32, source 0 = target; delays_s (S,2) seconds per source and microphone
    (`world_building.calculate_far_field_delays`).  -> mix (B,2,L), tgt (B,L), itf (B,L) float32, all divided by
    max|mix| + peak_eps per utterance (peak_eps=None: no division).
    Up to 4 sources of a length L = 4 * (2, 3, 5, 7-smooth) <= 102 400 are mixed by one cluster-resident kernel (spectra
    never leave the chip); anything else - or `multi_pass=True` - by the multi-pass transform through HBM."""
    io = _Io()
    src = io.take(sources, torch.float32)
    single = src.dim() == 2
    if single:
        src = src.unsqueeze(0)
    B, S, L = src.shape
    dl = np.ascontiguousarray(np.asarray(delays_s, dtype=np.float64).reshape(S, 2))
    lib = _lib.load()
    nbytes = lib.avz_farfield_mix_ws_bytes(B, S, L)
    if nbytes <= 0:
        raise _lib.AvzError(f"far_field_mix: unsupported shape B={B} S={S} L={L} (see include/avzoom.h)")
    ws = torch.empty(nbytes, dtype=torch.uint8, device=src.device)
    mix = torch.empty((B, 2, L), dtype=torch.float32, device=src.device)
    tgt = torch.empty((B, L), dtype=torch.float32, device=src.device)
    itf = torch.empty((B, L), dtype=torch.float32, device=src.device)
    fn = lib.avz_farfield_mix_passes_f32 if multi_pass else lib.avz_farfield_mix_f32
    _lib.check(fn(_ptr(src), dl.ctypes.data_as(C.POINTER(C.c_double)), B, S, L, float(fs),
                  -1.0 if peak_eps is None else float(peak_eps), _ptr(mix), _ptr(tgt), _ptr(itf), _ptr(ws), _stream()),
               "avz_farfield_mix_f32")
    if single:
        mix, tgt, itf = mix[0], tgt[0], itf[0]
    return io.give(mix), io.give(tgt), io.give(itf)


def fractional_delay(y, delay_sec: float, fs: float = 16000.0):
    """world_building.py:46-52 `apply_frac_delay`: whole-signal phase-ramp delay of (..., L) signals."""
    io = _Io()
    y = io.take(y, torch.float32)
    lead = y.shape[:-1]
    src = y.reshape(-1, 1, y.shape[-1])
    _, out, _ = far_field_mix(src, [[delay_sec, delay_sec]], fs, peak_eps=None)
    return io.give(out.reshape(*lead, y.shape[-1]))


def pcm16_to_float(pcm: torch.Tensor) -> torch.Tensor:
    """int16 CUDA tensor -> float32 = pcm / 32768 (what `soundfile.read(dtype='float32')` returns; oracle_debug.py:35-39)."""
    if pcm.dtype != torch.int16 or not pcm.is_cuda:
        raise _lib.AvzError("pcm16_to_float needs an int16 CUDA tensor")
    pcm = pcm.contiguous()
    out = torch.empty(pcm.shape, dtype=torch.float32, device=pcm.device)
    _lib.check(_lib.load().avz_pcm16_to_f32(_ptr(pcm), pcm.numel(), _ptr(out), _stream()), "avz_pcm16_to_f32")
    return out


def float_to_pcm16(x: torch.Tensor) -> torch.Tensor:
    """float32 CUDA tensor -> int16 as `soundfile.write` stores PCM_16 (round(x * 32767), clipped; oracle_debug.py:96)."""
    if x.dtype != torch.float32 or not x.is_cuda:
        raise _lib.AvzError("float_to_pcm16 needs a float32 CUDA tensor")
    x = x.contiguous()
    out = torch.empty(x.shape, dtype=torch.int16, device=x.device)
    _lib.check(_lib.load().avz_f32_to_pcm16(_ptr(x), x.numel(), _ptr(out), _stream()), "avz_f32_to_pcm16")
    return out


# ------------------------------------------------------------------------------------------ fused passes
def alloc_kept_spectrum(mix: torch.Tensor, cfg: MvdrConfig, ibm: bool = False) -> Optional[torch.Tensor]:
    """Workspace in which pass A keeps the mix spectrum for pass B (fast path shapes only, else None).
    `ibm`: for the oracle-IBM pass A, which keeps a spectrum at n_fft 512 only (the learned-mask pass A also does at
    n_fft 1024 / hop 512)."""
    B, _, L = mix.shape
    if ibm and cfg.n_fft != 512:
        return None
    n = _lib.load().avz_spec_ws_bytes(B, L, cfg.n_fft, cfg.hop)
    if n <= 0:
        return None
    # ~34 bytes per input sample (n_fft 512; 16 at n_fft 1024).  If it does not fit, pass B recomputes the transform instead (bit-identical result,
    # ~8 % slower): huge batches degrade gracefully instead of running out of memory.  (No cudaMemGetInfo here: it
    # costs milliseconds per call and does not see the blocks torch's caching allocator can reuse.)
    try:
        return torch.empty((int(n),), dtype=torch.uint8, device=mix.device)
    except torch.cuda.OutOfMemoryError:
        return None


def ibm_covariance(mix: torch.Tensor, tgt: torch.Tensor, itf: torch.Tensor, cfg: MvdrConfig,
                   spec: Optional[torch.Tensor] = None, sparse: bool = False, postmask: bool = False):
    """Pass A (oracle_debug.py:42-64) without storing any spectrum - or, with `spec` (alloc_kept_spectrum), keeping
    the packed mix spectrum so that pass B can skip its forward transform.
    mix [B,2,L], tgt [B,L], itf [B,L] -> (ibm_bits [B,T,ceil(F/32)] int32, R packed [B,F,4], msum [B,F])."""
    lib = _lib.load()
    B, _, L = mix.shape
    if tuple(tgt.shape) != (B, L) or tuple(itf.shape) != (B, L) or mix.shape[1] != 2:
        raise ValueError(f"mix {tuple(mix.shape)}, tgt {tuple(tgt.shape)}, itf {tuple(itf.shape)}: need (B,2,L), (B,L), (B,L)")
    F, T = cfg.n_freq, num_frames(L, cfg.n_fft, cfg.hop)
    dev = mix.device
    bits = torch.empty((B, T, (F + 31) // 32), dtype=torch.int32, device=dev)
    Rp = torch.empty((B, F, 4), dtype=torch.float32, device=dev)
    ms = torch.empty((B, F), dtype=torch.float32, device=dev)
    nws = lib.avz_ibm_cov_ws_bytes(B, L, cfg.n_fft, cfg.hop)
    if nws < 0:
        _lib.check(-1, "avz_ibm_cov_ws_bytes")
    ws = torch.empty((max(int(nws), 4),), dtype=torch.uint8, device=dev)
    if spec is not None and sparse:
        # only the bins the post-filter 1 - noise mask lets through are kept (read back with mvdr_apply(..., sparse=True))
        _lib.check(lib.avz_ibm_cov_keep_sparse_f32(_ptr(mix), _ptr(tgt), _ptr(itf), B, L, cfg.n_fft, cfg.hop,
                                                   float(cfg.norm_eps), _ptr(bits), _ptr(Rp), _ptr(ms), _ptr(ws), _ptr(spec),
                                                   _stream()), "avz_ibm_cov_keep_sparse_f32")
    elif spec is not None and postmask:
        # pass B will apply 1 - noise mask with these bits: sectors it zeroes anyway are not written (mvdr_apply as usual);
        # `spec` must have been zeroed once (pipeline.OracleMvdr does) - see avz_ibm_cov_keep_postmask_f32
        _lib.check(lib.avz_ibm_cov_keep_postmask_f32(_ptr(mix), _ptr(tgt), _ptr(itf), B, L, cfg.n_fft, cfg.hop,
                                                     float(cfg.norm_eps), _ptr(bits), _ptr(Rp), _ptr(ms), _ptr(ws), _ptr(spec),
                                                     _stream()), "avz_ibm_cov_keep_postmask_f32")
    elif spec is not None:
        _lib.check(lib.avz_ibm_cov_keep_f32(_ptr(mix), _ptr(tgt), _ptr(itf), B, L, cfg.n_fft, cfg.hop,
                                            float(cfg.norm_eps), _ptr(bits), _ptr(Rp), _ptr(ms), _ptr(ws), _ptr(spec),
                                            _stream()), "avz_ibm_cov_keep_f32")
    else:
        _lib.check(lib.avz_ibm_cov_f32(_ptr(mix), _ptr(tgt), _ptr(itf), B, L, cfg.n_fft, cfg.hop, float(cfg.norm_eps),
                                       _ptr(bits), _ptr(Rp), _ptr(ms), _ptr(ws), _stream()), "avz_ibm_cov_f32")
    return bits, Rp, ms


def ibm_exact_bits(tgt: torch.Tensor, itf: torch.Tensor, cfg: MvdrConfig) -> torch.Tensor:
    """The IBM of oracle_debug.py:42-53 with every bin decided in float64 on the GPU (slow reference form).
    tgt, itf [B,L] -> ibm_bits [B,T,9] int32."""
    B, L = tgt.shape
    T = num_frames(L, cfg.n_fft, cfg.hop)
    bits = torch.empty((B, T, (cfg.n_freq + 31) // 32), dtype=torch.int32, device=tgt.device)
    ws = torch.empty((16,), dtype=torch.uint8, device=tgt.device)
    _lib.check(_lib.load().avz_ibm_exact_f32(_ptr(tgt), _ptr(itf), B, L, cfg.n_fft, cfg.hop, _ptr(bits), _ptr(ws),
                                             _stream()), "avz_ibm_exact_f32")
    return bits


def wave_masked_covariance(mix: torch.Tensor, mask: torch.Tensor, cfg: MvdrConfig,
                           spec: Optional[torch.Tensor] = None, wide: bool = False):
    """Learned-mask pass A (full_audio.../inference.py:90,102-108): mix [B,2,L], target-probability mask [B,F,T]
    -> (R packed [B,F,4], msum [B,F]); noise weight = 1 - mask.  `wide`: STFT and accumulation in float64 from the
    same float32 inputs (R, msum float64) for ill-conditioned consumers; no kept spectrum in that mode."""
    lib = _lib.load()
    B, _, L = mix.shape
    F = cfg.n_freq
    dev = mix.device
    T = num_frames(L, cfg.n_fft, cfg.hop)
    if tuple(mask.shape) != (B, F, T) or mask.dtype != torch.float32 or not mask.is_contiguous():
        raise ValueError(f"mask must be a contiguous float32 tensor of shape {(B, F, T)}, got {tuple(mask.shape)} {mask.dtype}")
    if wide:
        Rp = torch.empty((B, F, 4), dtype=torch.float64, device=dev)
        ms = torch.empty((B, F), dtype=torch.float64, device=dev)
        ws = torch.empty((int(lib.avz_wave_mask_cov_f64_ws_bytes(B, L, cfg.n_fft, cfg.hop)),), dtype=torch.uint8, device=dev)
        _lib.check(lib.avz_wave_mask_cov_f64(_ptr(mix), _ptr(mask), B, L, cfg.n_fft, cfg.hop, float(cfg.sqrt_eps),
                                             float(cfg.norm_eps), _ptr(Rp), _ptr(ms), _ptr(ws), _stream()),
                   "avz_wave_mask_cov_f64")
        return Rp, ms
    Rp = torch.empty((B, F, 4), dtype=torch.float32, device=dev)
    ms = torch.empty((B, F), dtype=torch.float32, device=dev)
    nws = lib.avz_ibm_cov_ws_bytes(B, L, cfg.n_fft, cfg.hop)
    ws = torch.empty((max(int(nws), 4),), dtype=torch.uint8, device=dev)
    if spec is not None:
        _lib.check(lib.avz_wave_mask_cov_keep_f32(_ptr(mix), _ptr(mask), B, L, cfg.n_fft, cfg.hop, float(cfg.sqrt_eps),
                                                  float(cfg.norm_eps), _ptr(Rp), _ptr(ms), _ptr(ws), _ptr(spec),
                                                  _stream()), "avz_wave_mask_cov_keep_f32")
    else:
        _lib.check(lib.avz_wave_mask_cov_f32(_ptr(mix), _ptr(mask), B, L, cfg.n_fft, cfg.hop, float(cfg.sqrt_eps),
                                             float(cfg.norm_eps), _ptr(Rp), _ptr(ms), _ptr(ws), _stream()),
                   "avz_wave_mask_cov_f32")
    return Rp, ms


def mvdr_apply(mix: torch.Tensor, w: torch.Tensor, cfg: MvdrConfig, ibm_bits: Optional[torch.Tensor] = None,
               mask: Optional[torch.Tensor] = None, spec: Optional[torch.Tensor] = None,
               mask_staged: bool = False, sparse: bool = False) -> Tuple[torch.Tensor, torch.Tensor]:
    """Pass B (oracle_debug.py:80-93): STFT(mix) (or the spectrum pass A kept in `spec`) -> w^H y -> post-filter ->
    iSTFT/OLA.  -> (out [B,(T-1)*hop] un-normalised, peak [B]).
    `mask_staged`: `mask` is the one `wave_masked_covariance(..., spec)` has just been given - at n_fft 512 pass B then
    reuses the transposed copy that call left in `spec` instead of re-laying the mask a second time."""
    B, _, L = mix.shape
    T = num_frames(L, cfg.n_fft, cfg.hop)
    F = cfg.n_freq
    if tuple(w.shape) != (B, F, 2) or w.dtype != torch.complex64 or not w.is_contiguous():
        raise ValueError(f"w must be a contiguous complex64 tensor of shape {(B, F, 2)}, got {tuple(w.shape)} {w.dtype}")
    if ibm_bits is not None and (tuple(ibm_bits.shape) != (B, T, (F + 31) // 32) or ibm_bits.dtype != torch.int32):
        raise ValueError(f"ibm_bits must be int32 of shape {(B, T, (F + 31) // 32)}, got {tuple(ibm_bits.shape)}")
    if mask is not None and (tuple(mask.shape) != (B, F, T) or mask.dtype != torch.float32 or not mask.is_contiguous()):
        raise ValueError(f"mask must be a contiguous float32 tensor of shape {(B, F, T)}, got {tuple(mask.shape)}")
    if mask_staged and spec is not None and cfg.n_fft == 512 and cfg.post in ("floor", "mask"):
        mask = None
    out = torch.empty((B, (T - 1) * cfg.hop), dtype=torch.float32, device=mix.device)
    peak = torch.zeros((B,), dtype=torch.float32, device=mix.device)
    cc = cfg.to_c()
    if spec is not None and sparse:
        _lib.check(_lib.load().avz_mvdr_apply_kept_sparse_f32(_ptr(spec), _ptr(w), _ptr(ibm_bits), B, L, cfg.n_fft, cfg.hop,
                                                              C.byref(cc), _ptr(out), _ptr(peak), _stream()),
                   "avz_mvdr_apply_kept_sparse_f32")
    elif spec is not None:
        _lib.check(_lib.load().avz_mvdr_apply_kept_f32(_ptr(spec), _ptr(w), _ptr(ibm_bits), _ptr(mask), B, L, cfg.n_fft,
                                                       cfg.hop, C.byref(cc), _ptr(out), _ptr(peak), _stream()),
                   "avz_mvdr_apply_kept_f32")
    else:
        _lib.check(_lib.load().avz_mvdr_apply_f32(_ptr(mix), _ptr(w), _ptr(ibm_bits), _ptr(mask), B, L, cfg.n_fft,
                                                  cfg.hop, C.byref(cc), _ptr(out), _ptr(peak), _stream()),
                   "avz_mvdr_apply_f32")
    return out, peak


def _batchify(mix, tgt, itf, io: _Io):
    mix = io.take(mix, torch.float32)
    single = mix.dim() == 2
    if single:
        mix = mix[None]
    tgt = None if tgt is None else io.take(tgt, torch.float32).reshape(mix.shape[0], -1)
    itf = None if itf is None else io.take(itf, torch.float32).reshape(mix.shape[0], -1)
    return mix.contiguous(), tgt, itf, single


def oracle_mask_mvdr(mix, tgt, itf, cfg: MvdrConfig = PRESETS["baseline_oracle"], return_parts: bool = False):
    """Oracle IBM mask-MVDR end to end (oracle_debug.py:42-94), fused into two passes over the waveforms.
    mix (B,2,L) or (2,L); tgt, itf (B,L) or (L,) -> waveform (B,(T-1)*hop) float32 (peak-normalised per
    utterance when cfg.peak_eps is not None)."""
    io = _Io()
    mix, tgt, itf, single = _batchify(mix, tgt, itf, io)
    spec = alloc_kept_spectrum(mix, cfg, ibm=True)
    sparse = False      # the sparse kept spectrum saves 60 % of the traffic but costs more than it saves (pipeline.OracleMvdr)
    bits, Rp, ms = ibm_covariance(mix, tgt, itf, cfg, spec, sparse=sparse)   # (spec is fresh, uninitialised memory: no postmask)
    w = mvdr_weights(Rp, steering_vectors(cfg, mix.device), cfg)
    out, peak = mvdr_apply(mix, w, cfg, ibm_bits=bits if cfg.post == "one_minus_noise" else None, spec=spec, sparse=sparse)
    parts = None
    if return_parts:
        parts = {"ibm_bits": bits, "R": Rp, "msum": ms, "w": w, "x_raw": out.clone(), "peak": peak}
    if cfg.peak_eps is not None:
        peak_normalise(out, peak, cfg.peak_eps)
    res = io.give(out[0] if single else out)
    return (res, parts) if return_parts else res


def learned_mask_mvdr(mix, mask, cfg: MvdrConfig = PRESETS["baseline_learned"]):
    """Learned-mask MVDR given the mask (process_chunk after the model call,
    full_audio.../inference.py:99-117).  mix (B,2,L), mask (B,F,T) target probability -> (B,(T-1)*hop)."""
    io = _Io()
    mix, _, _, single = _batchify(mix, None, None, io)
    mask = io.take(mask, torch.float32).reshape(mix.shape[0], cfg.n_freq, -1).contiguous()
    spec = alloc_kept_spectrum(mix, cfg)
    Rp, _ = wave_masked_covariance(mix, mask, cfg, spec)
    w = mvdr_weights(Rp, steering_vectors(cfg, mix.device), cfg)
    out, peak = mvdr_apply(mix, w, cfg, mask=mask if cfg.post in ("floor", "mask") else None, spec=spec, mask_staged=True)
    if cfg.peak_eps is not None:
        peak_normalise(out, peak, cfg.peak_eps)
    return io.give(out[0] if single else out)

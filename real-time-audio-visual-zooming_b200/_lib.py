"""ctypes binding of libavzoom.so (the C ABI declared in include/avzoom.h).

There is no fallback: if the shared library is missing or a call fails, an exception is raised.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("AVZ_LIB") or os.path.join(_HERE, "libavzoom.so")  # AVZ_LIB: A/B builds of the same ABI

_lib = None


class AvzError(RuntimeError):
    pass


class AvzMvdrCfg(C.Structure):
    """Mirror of `struct AvzMvdrCfg` (include/avzoom.h)."""
    _fields_ = [
        ("sigma", C.c_float),
        ("norm_eps", C.c_float),
        ("sqrt_eps", C.c_float),
        ("w_eps", C.c_float),
        ("hp_bins", C.c_int32),
        ("hp_mode", C.c_int32),
        ("post_mode", C.c_int32),
        ("post_floor", C.c_float),
    ]


class AvzChunkView(C.Structure):
    """Mirror of `struct AvzChunkView` (include/avzoom.h)."""
    _fields_ = [("rec_len", C.c_int64), ("n_windows", C.c_int32), ("stride", C.c_int32)]


POST_NONE, POST_ONE_MINUS_NOISE, POST_FLOOR, POST_MASK = 0, 1, 2, 3
HP_NONE, HP_ZERO, HP_MIC0 = 0, 1, 2
FEAT_LOGMAG_IPD, FEAT_LOGMAG_IPD_WRAPPED, FEAT_PHYSICS_NHWC = 0, 1, 2

_p = C.c_void_p
_i = C.c_int
_l = C.c_int64
_f = C.c_float
_d = C.c_double

# name -> (restype, argtypes); every symbol include/avzoom.h declares
SIGNATURES = {
    "avz_version": (_i, []),
    "avz_last_error": (C.c_char_p, []),
    "avz_build_info": (C.c_char_p, []),
    "avz_init": (_i, [_i]),
    "avz_profile_enable": (_i, [_i]),
    "avz_profile_get": (_i, [C.POINTER(C.c_float), _i]),
    "avz_num_frames": (_l, [_l, _i, _i]),
    "avz_stft_f32": (_i, [_p, _i, _i, _l, _i, _i, _p, _p]),
    "avz_istft_f32": (_i, [_p, _i, _i, _i, _i, _p, _p, _p]),
    "avz_peak_normalise_f32": (_i, [_p, _i, _l, _p, _f, _p]),
    "avz_ibm_cov_ws_bytes": (_l, [_i, _l, _i, _i]),
    "avz_ibm_cov_f32": (_i, [_p, _p, _p, _i, _l, _i, _i, _f, _p, _p, _p, _p, _p]),
    "avz_ibm_exact_f32": (_i, [_p, _p, _i, _l, _i, _i, _p, _p, _p]),
    "avz_wave_mask_cov_f32": (_i, [_p, _p, _i, _l, _i, _i, _f, _f, _p, _p, _p, _p]),
    "avz_spec_mask_cov_f32": (_i, [_p, _p, _i, _i, _i, _f, _f, _p, _p, _p]),
    "avz_mvdr_weights_f32": (_i, [_p, _p, _i, _i, C.POINTER(AvzMvdrCfg), _p, _p]),
    "avz_hybrid_null_weights_f32": (_i, [_p, _p, _i, _i, _i, _i, _p, _p]),
    "avz_beamform_f32": (_i, [_p, _p, _i, _i, _i, _p, _p]),
    "avz_mvdr_apply_f32": (_i, [_p, _p, _p, _p, _i, _l, _i, _i, C.POINTER(AvzMvdrCfg), _p, _p, _p]),
    "avz_spec_ws_bytes": (_l, [_i, _l, _i, _i]),
    "avz_ibm_cov_keep_f32": (_i, [_p, _p, _p, _i, _l, _i, _i, _f, _p, _p, _p, _p, _p, _p]),
    "avz_wave_mask_cov_keep_f32": (_i, [_p, _p, _i, _l, _i, _i, _f, _f, _p, _p, _p, _p, _p]),
    "avz_mvdr_apply_kept_f32": (_i, [_p, _p, _p, _p, _i, _l, _i, _i, C.POINTER(AvzMvdrCfg), _p, _p, _p]),
    "avz_mvdr_apply_kept_norm_f32": (_i, [_p, _p, _p, _p, _i, _l, _i, _i, C.POINTER(AvzMvdrCfg), _f, _p, _p, _p]),
    "avz_ibm_cov_weights_keep_f32": (_i, [_p, _p, _p, _i, _l, _i, _i, C.POINTER(AvzMvdrCfg), _p, _p, _p, _p, _p, _p, _p, _i, _p]),
    "avz_ibm_cov_keep_postmask_f32": (_i, [_p, _p, _p, _i, _l, _i, _i, _f, _p, _p, _p, _p, _p, _p]),
    "avz_ibm_cov_keep_sparse_f32": (_i, [_p, _p, _p, _i, _l, _i, _i, _f, _p, _p, _p, _p, _p, _p]),
    "avz_mvdr_apply_kept_sparse_f32": (_i, [_p, _p, _p, _i, _l, _i, _i, C.POINTER(AvzMvdrCfg), _p, _p, _p]),
    "avz_oracle_fused_ws_bytes": (_l, [_i, _l, _i, _i]),
    "avz_oracle_fused_f32": (_i, [_p, _p, _p, _i, _l, _i, _i, C.POINTER(AvzMvdrCfg), _f, _p, _p, _p, _p, _p, _p, _p, _p, _p]),
    "avz_stream_state_bytes": (_l, [_i]),
    "avz_stream_step_f32": (_i, [_p, _p, _p, _p, _i, _i, _i, _f, C.POINTER(AvzMvdrCfg), _p, _p]),
    "avz_mag_greater_f32": (_i, [_p, _p, _l, _p, _p]),
    "avz_irm_f32": (_i, [_p, _p, _l, _p, _p]),
    "avz_geometric_mask_f32": (_i, [_p, _i, _i, _i, _p, _p]),
    "avz_ibm_unpack_f32": (_i, [_p, _i, _i, _i, _p, _p]),
    "avz_features_f32": (_i, [_p, _i, _i, _i, _i, _p, _p]),
    "avz_wave_features_f32": (_i, [_p, _i, _l, _i, _i, _i, _p, _p]),
    "avz_sir_f32": (_i, [_p, _p, _p, _i, _l, _l, _p, _p]),
    "avz_farfield_mix_ws_bytes": (_l, [_i, _i, _l]),
    "avz_farfield_mix_f32": (_i, [_p, C.POINTER(C.c_double), _i, _i, _l, C.c_double, _f, _p, _p, _p, _p, _p]),
    "avz_farfield_mix_passes_f32": (_i, [_p, C.POINTER(C.c_double), _i, _i, _l, C.c_double, _f, _p, _p, _p, _p, _p]),
    "avz_chunk_features_f32": (_i, [_p, _i, C.POINTER(AvzChunkView), _l, _i, _i, _i, _p, _p]),
    "avz_chunk_mask_cov_f32": (_i, [_p, _p, _i, C.POINTER(AvzChunkView), _l, _i, _i, _f, _f, _p, _p, _p, _p, _p]),
    "avz_chunk_mvdr_apply_f32": (_i, [_p, _p, _p, _p, _i, C.POINTER(AvzChunkView), _l, _i, _i, C.POINTER(AvzMvdrCfg), _p, _p, _p]),
    "avz_chunk_ola_f32": (_i, [_p, _i, _i, _l, _l, _i, _l, _p, _p, _p]),
    "avz_pcm16_frames_to_planar_f32": (_i, [_p, _i, _l, _i, _p, _p]),
    "avz_stft_f64": (_i, [_p, _i, _i, _l, _i, _i, _p, _p]),
    "avz_istft_f64_ws_bytes": (_l, [_i, _i, _i]),
    "avz_istft_f64": (_i, [_p, _i, _i, _i, _i, _p, _p, _p]),
    "avz_peak_normalise_f64": (_i, [_p, _i, _l, _d, _p, _p]),
    "avz_geometric_mask_f64": (_i, [_p, _i, _i, _i, _p, _p]),
    "avz_spec_mask_cov_f64": (_i, [_p, _p, _i, _i, _i, _d, _d, _p, _p, _p]),
    "avz_wave_mask_cov_f64_ws_bytes": (_l, [_i, _l, _i, _i]),
    "avz_wave_mask_cov_f64": (_i, [_p, _p, _i, _l, _i, _i, _d, _d, _p, _p, _p, _p]),
    "avz_chunk_mask_cov_f64": (_i, [_p, _p, _i, C.POINTER(AvzChunkView), _l, _i, _i, _d, _d, _p, _p, _p, _p]),
    "avz_hybrid_null_weights_f64_w32": (_i, [_p, _p, _i, _i, _i, _i, _p, _p]),
    "avz_mvdr_weights_f64": (_i, [_p, _p, _i, _i, _d, _d, _i, _i, _p, _p]),
    "avz_hybrid_null_weights_f64": (_i, [_p, _p, _i, _i, _i, _i, _p, _p]),
    "avz_beamform_f64": (_i, [_p, _p, _i, _i, _i, _p, _p]),
    "avz_pcm16_to_f32": (_i, [_p, _l, _p, _p]),
    "avz_f32_to_pcm16": (_i, [_p, _l, _p, _p]),
}


def load():
    """Load libavzoom.so once; raise loudly if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise AvzError(
            f"{LIB_PATH} is missing: build it with `python -c \"import __graft_entry__ as g; g.build()\"` "
            "(or `make -C real-time-audio-visual-zooming_b200/csrc`). There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the library is stale
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int, what: str):
    if rc != 0:
        msg = load().avz_last_error()
        raise AvzError(f"{what} failed with code {rc}: {msg.decode() if msg else ''}")

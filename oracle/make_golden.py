#!/usr/bin/env python
"""Generate tests/golden/*.npz by importing and running the UNMODIFIED reference code.

Run in the authoring container only (needs /root/reference; the GPU box does not have it):

    python oracle/make_golden.py

Recipe: the reference's absent I/O dependencies (soundfile, matplotlib,
tensorflow, ...) are replaced by empty stub modules; `soundfile.read/write` are backed by an
in-memory table so the reference's `main()` functions run end to end without touching disk
formats we cannot read; everything numerical is the reference's own code on this image's
numpy/scipy/torch.  Outputs are small (seeded random inputs, a 2 s excerpt of the reference's
real-speech fixture) and are committed together with this script.
"""
from __future__ import annotations

import contextlib
import importlib.util
import io
import os
import sys
import tempfile
import types
import wave

import numpy as np

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")


# ------------------------------------------------------------------ stubs
class _SoundFileShim(types.ModuleType):
    """soundfile stand-in: PCM16 WAV through stdlib `wave`; also records what was written and can
    be told to hand back float64 regardless of the requested dtype."""

    def __init__(self):
        super().__init__("soundfile")
        self.written = {}
        self.force_float64 = False

    def read(self, path, dtype="float64"):
        with wave.open(path, "rb") as w:
            n, ch, fs = w.getnframes(), w.getnchannels(), w.getframerate()
            assert w.getsampwidth() == 2
            pcm = np.frombuffer(w.readframes(n), dtype="<i2")
        data = pcm.astype(np.float64) / 32768.0
        if ch > 1:
            data = data.reshape(n, ch)
        if not self.force_float64:
            data = data.astype(dtype)
        return data, fs

    def write(self, path, data, fs):
        self.written[os.path.basename(path)] = np.array(data, dtype=np.float64, copy=True)


def write_wav_pcm16(path, pcm: np.ndarray, fs=16000):
    pcm = np.asarray(pcm, dtype="<i2")
    ch = 1 if pcm.ndim == 1 else pcm.shape[1]
    with wave.open(path, "wb") as w:
        w.setnchannels(ch)
        w.setsampwidth(2)
        w.setframerate(fs)
        w.writeframes(pcm.tobytes())


def read_wav_pcm16(path):
    with wave.open(path, "rb") as w:
        n, ch = w.getnframes(), w.getnchannels()
        pcm = np.frombuffer(w.readframes(n), dtype="<i2")
    return pcm.reshape(n, ch) if ch > 1 else pcm


def install_stubs():
    sf = _SoundFileShim()
    sys.modules["soundfile"] = sf
    mpl = types.ModuleType("matplotlib")
    plt = types.ModuleType("matplotlib.pyplot")
    for name in ("figure", "imshow", "title", "savefig", "show", "plot", "close", "colorbar"):
        setattr(plt, name, lambda *a, **k: None)
    mpl.pyplot = plt
    mpl.use = lambda *a, **k: None
    sys.modules["matplotlib"] = mpl
    sys.modules["matplotlib.pyplot"] = plt
    for name in ("tensorflow", "kagglehub", "librosa", "pyroomacoustics", "nara_wpe", "nara_wpe.wpe",
                 "nara_wpe.utils", "mir_eval", "mir_eval.separation", "pystoi", "pesq"):
        sys.modules[name] = types.ModuleType(name)
    sys.modules["mir_eval.separation"].bss_eval_sources = None
    sys.modules["mir_eval"].separation = sys.modules["mir_eval.separation"]
    sys.modules["pystoi"].stoi = None
    sys.modules["pesq"].pesq = None
    return sf


def load_file_module(name, path, cwd):
    """Import a reference script by path with cwd set to its directory (they read config.json
    relative to cwd at import time), silencing its prints."""
    old = os.getcwd()
    os.chdir(cwd)
    try:
        spec = importlib.util.spec_from_file_location(name, path)
        mod = importlib.util.module_from_spec(spec)
        with contextlib.redirect_stdout(io.StringIO()):
            spec.loader.exec_module(mod)
    finally:
        os.chdir(old)
    return mod


def main():
    os.makedirs(OUT, exist_ok=True)
    sf = install_stubs()
    sys.path.insert(0, REF)
    sys.path.insert(0, os.path.join(REF, "Final_pipeline"))
    import rt_av_zoom.core.masked_mvdr as ref_mm
    import rt_av_zoom.core.oracle_debug as ref_od
    with contextlib.redirect_stdout(io.StringIO()):
        import src.inference as ref_fp_inf
        import src.metrics as ref_fp_met
    ref_tfl = load_file_module("ref_tfl_inf", os.path.join(REF, "rt_av_zoom/core/tf_lite_version/inference.py"),
                               os.path.join(REF, "rt_av_zoom/core/tf_lite_version"))
    ref_wb = load_file_module("ref_wb", os.path.join(REF, "rt_av_zoom/core/tf_lite_version/world_building.py"),
                              os.path.join(REF, "rt_av_zoom/core/tf_lite_version"))
    ref_full = load_file_module("ref_full_inf",
                                os.path.join(REF, "rt_av_zoom/core/full_audio_generating_pipeline/inference.py"),
                                os.path.join(REF, "rt_av_zoom/core/full_audio_generating_pipeline"))
    ref_rm = load_file_module("ref_run_metrics", os.path.join(REF, "scripts/run_metrics.py"), REF)

    # ---------------------------------------------------------------- 4. far-field mixer
    def gen_mixer():
        """world_building.mix_and_save run unmodified on three PCM16 'files' (lengths the GPU mixer's two-factor
        transform supports: 8000 = 2^6 * 125), plus apply_frac_delay at L = 4000."""
        r4 = np.random.default_rng(20261019)
        m = {}
        L4 = 8000
        src_pcm = np.clip(np.rint(r4.standard_normal((3, L4)) * 6000.0), -32768, 32767).astype("<i2")
        m["src_pcm"] = src_pcm
        with tempfile.TemporaryDirectory() as td:
            old = os.getcwd()
            os.chdir(td)
            try:
                files = []
                for i in range(3):
                    fn = os.path.join(td, f"src{i}.wav")
                    write_wav_pcm16(fn, src_pcm[i])
                    files.append(fn)
                sf.written.clear()
                with contextlib.redirect_stdout(io.StringIO()):
                    ref_wb.mix_and_save(files, "g")
                m["mix"] = sf.written["mixture_g.wav"]            # (L, 2)
                m["tgt"] = sf.written["target_ref_g.wav"]
                m["itf"] = sf.written["interf_ref_g.wav"]
            finally:
                os.chdir(old)
        m["angles"] = np.array([ref_wb.ANGLE_TARGET, ref_wb.ANGLE_INTERFERER_A, ref_wb.ANGLE_INTERFERER_B], dtype=np.float64)
        m["d_c_fs"] = np.array([ref_wb.D, ref_wb.C, ref_wb.FS], dtype=np.float64)
        yd2 = r4.standard_normal(4000).astype(np.float32)
        m["fd2_in"] = yd2
        m["fd2_out"] = ref_wb.apply_frac_delay(yd2.astype(np.float64), -4.1e-5, 16000)
        np.savez_compressed(os.path.join(OUT, "ref_mixer.npz"), **m)

    if "--only-mixer" in sys.argv:
        gen_mixer()
        print("ref_mixer.npz", os.path.getsize(os.path.join(OUT, "ref_mixer.npz")))
        return

    rng = np.random.default_rng(20261018)

    # ---------------------------------------------------------------- 1. importable helpers
    g = {}
    sv_args = np.array([[90.0, 1000.0, 0.01, 343.0], [40.0, 3125.0, 0.04, 343.0], [130.0, 7968.75, 0.08, 340.0],
                        [0.0, 31.25, 0.04, 343.0], [65.0, 0.0, 0.04, 343.0]])
    g["sv_args"] = sv_args
    g["sv_out"] = np.stack([ref_mm.get_steering_vector(*a)[:, 0] for a in sv_args])
    g["sv_full_out"] = np.stack([ref_full.get_steering_vector(*a)[:, 0] for a in sv_args])
    f_bins = np.fft.rfftfreq(1024, 1 / 16000.0)
    g["asv_f_bins"] = f_bins
    g["asv_out"] = ref_tfl.get_all_steering_vectors(f_bins, 90.0, 0.04, 343.0)
    g["asv_out_40"] = ref_tfl.get_all_steering_vectors(f_bins, 40.0, 0.04, 343.0)
    g["constants_masked_mvdr"] = np.array([ref_mm.FS, ref_mm.D, ref_mm.C, ref_mm.ANGLE_TARGET, ref_mm.N_MICS,
                                           ref_mm.SIGMA, ref_mm.N_FFT, ref_mm.N_HOP], dtype=np.float64)

    Fg, Tg = 129, 40
    Yg = (rng.standard_normal((2, Fg, Tg)) + 1j * rng.standard_normal((2, Fg, Tg)))
    Yg[1, :, :5] = Yg[0, :, :5] * 2.0          # identical phase -> mask 0.01 there (exact tie path)
    Yg[:, 3, 7] = 0.0
    Yg = Yg.astype(np.complex64).astype(np.complex128)   # exactly float32-representable inputs
    g["geo_Y"] = Yg.astype(np.complex64)
    g["geo_mask"] = ref_mm.compute_hard_geometric_mask(Yg, None)

    Fb, Tb = 513, 64
    Yb = (rng.standard_normal((2, Fb, Tb)) + 1j * rng.standard_normal((2, Fb, Tb))) * 0.01
    maskb = rng.random((Fb, Tb))
    maskb[5, :] = 1.0                            # empty noise mask in one bin
    Yb = Yb.astype(np.complex64).astype(np.complex128)
    maskb = maskb.astype(np.float32).astype(np.float64)
    dvb = ref_tfl.get_all_steering_vectors(f_bins, 90.0, 0.04, 343.0)
    g["bm_Y"] = Yb.astype(np.complex64)
    g["bm_mask"] = maskb.astype(np.float32)
    g["bm_out"] = ref_tfl.batch_mvdr(Yb, maskb, f_bins, dvb, 1e-5)
    dvb40 = ref_tfl.get_all_steering_vectors(f_bins, 40.0, 0.04, 343.0)
    g["bm_out_40"] = ref_tfl.batch_mvdr(Yb, maskb, f_bins, dvb40, 1e-3)

    g["hn_out"] = ref_fp_inf.hybrid_hard_null_bf(Yb, maskb, f_bins)

    sig = rng.standard_normal((3, 6000)).astype(np.float32).astype(np.float64)
    est = 0.8 * sig[0] + 0.1 * sig[1] + 0.05 * sig[2]
    est = est.astype(np.float32).astype(np.float64)
    g["score_in"] = np.stack([est, sig[0], sig[1]]).astype(np.float32)
    g["score_osinr_osir"] = np.array(ref_fp_met.calculate_osnr_osir(est, sig[0], sig[1]))
    g["score_sdr_sir"] = np.array(ref_rm.calculate_metrics_manual(est, sig[0], sig[1]))
    g["score_full_manual"] = np.array(ref_full.calculate_metrics_manual(est, sig[0], sig[1]))

    g["ffd_angles"] = np.array([90.0, 40.0, 130.0, 65.0, 155.0, 20.0])
    g["ffd_out"] = np.array([ref_wb.calculate_far_field_delays(a, 0.04, 343.0) for a in g["ffd_angles"]])
    yd = rng.standard_normal(4001).astype(np.float32).astype(np.float64)
    g["fd_in"] = yd
    g["fd_out"] = ref_wb.apply_frac_delay(yd, 3.3e-5, 16000)
    np.savez_compressed(os.path.join(OUT, "ref_helpers.npz"), **g)

    # ---------------------------------------------------------------- 2. real-speech excerpt
    lo, hi = 16000, 56000
    mix_pcm = read_wav_pcm16(os.path.join(REF, "data/inputs/mixture_3_sources_2.wav"))[lo:hi]
    tgt_pcm = read_wav_pcm16(os.path.join(REF, "data/inputs/target_reference_2.wav"))[lo:hi]
    int_pcm = read_wav_pcm16(os.path.join(REF, "data/inputs/interference_reference_2.wav"))[lo:hi]
    assert mix_pcm.shape == (hi - lo, 2) and tgt_pcm.shape == (hi - lo,) and int_pcm.shape == (hi - lo,)

    e = {"mix_pcm": mix_pcm, "tgt_pcm": tgt_pcm, "int_pcm": int_pcm}
    with tempfile.TemporaryDirectory() as td:
        old = os.getcwd()
        os.chdir(td)
        try:
            # oracle_debug.main(): hard-coded OUTDIR relative to cwd (oracle_debug.py:25,31-39)
            od_dir = os.path.join(td, ref_od.OUTDIR)
            os.makedirs(od_dir)
            write_wav_pcm16(os.path.join(od_dir, "mixture.wav"), mix_pcm)
            write_wav_pcm16(os.path.join(od_dir, "target_reference.wav"), tgt_pcm)
            write_wav_pcm16(os.path.join(od_dir, "interference_reference.wav"), int_pcm)
            for force64, key in ((False, "oracle_debug_out_f32read"), (True, "oracle_debug_out_f64read")):
                sf.force_float64 = force64
                sf.written.clear()
                with contextlib.redirect_stdout(io.StringIO()):
                    ref_od.main()
                e[key] = sf.written["output_oracle.wav"]
            # masked_mvdr.main(dir): expects <dir>/mixture_3_sources.wav (masked_mvdr.py:59)
            world = os.path.join(td, "run", "World_Outputs")
            os.makedirs(world)
            write_wav_pcm16(os.path.join(world, "mixture_3_sources.wav"), mix_pcm)
            for force64, key in ((False, "masked_mvdr_out_f32read"), (True, "masked_mvdr_out_f64read")):
                sf.force_float64 = force64
                sf.written.clear()
                with contextlib.redirect_stdout(io.StringIO()):
                    ref_mm.main(world)
                e[key] = sf.written["output_masked_mvdr.wav"]
            sf.force_float64 = False
        finally:
            os.chdir(old)
    np.savez_compressed(os.path.join(OUT, "ref_speech_excerpt.npz"), **e)

    # ---------------------------------------------------------------- 3. learned-mask chunk path
    import torch
    torch.manual_seed(0)
    model = ref_full.FreqPreservingUNet()
    model.eval()
    c = {}
    L3 = 40000                                   # 2.5 s -> ceil(40000/16000) = 3 windows
    mix3 = (mix_pcm[:L3].astype(np.float64) / 32768.0)
    chunk0 = mix3[:32000].astype(np.float32)
    with contextlib.redirect_stdout(io.StringIO()):
        c["chunk0_out"] = ref_full.process_chunk(chunk0, model)
    c["state_keys"] = np.array(list(model.state_dict().keys()))
    # masks the model produces for each of main_deploy's windows (float32 reads, as the reference does)
    import scipy.signal
    masks = []
    for i in range(int(np.ceil(L3 / 16000))):
        ch = mix3[i * 16000:i * 16000 + 32000].astype(np.float32)
        if len(ch) < 32000:
            ch = np.pad(ch, ((0, 32000 - len(ch)), (0, 0)))
        _, _, Y = scipy.signal.stft(ch.T, fs=16000, nperseg=1024, noverlap=512)
        X = torch.from_numpy(np.stack([np.log(np.abs(Y[0]) + 1e-7), np.angle(Y[0]) - np.angle(Y[1])], 0)).float()[None]
        with torch.no_grad():
            masks.append(model(X)[0].numpy())
    c["masks"] = np.stack(masks).astype(np.float32)
    with tempfile.TemporaryDirectory() as td:
        old = os.getcwd()
        os.chdir(td)
        try:
            torch.save(model.state_dict(), "mask_3.pth")
            write_wav_pcm16("speech_TEST.wav", mix_pcm[:L3])
            sf.written.clear()
            with contextlib.redirect_stdout(io.StringIO()):
                ref_full.main_deploy("speech_TEST.wav")
            c["main_deploy_out"] = sf.written["enhanced_speech_TEST.wav"]
        finally:
            os.chdir(old)
    c["L"] = np.array(L3)
    np.savez_compressed(os.path.join(OUT, "ref_learned_chunk.npz"), **c)

    # ---------------------------------------------------------------- 5. the TFLite-era chunk drivers (row 9b) and NaN bins
    # process_audio_file (tf_lite_version/inference.py:245-391) and enhance_audio (Final_pipeline/src/inference.py:144-238)
    # run unmodified; only their TFLiteBeamformer (a tf.lite interpreter over a blob that is not shipped) is replaced by
    # an object that replays the masks of section 3, window by window.
    class ReplayBeamformer:
        def __init__(self, *a, **k):
            self.i = 0

        def predict_mask(self, log_mag, ipd):
            m = c["masks"][self.i]
            self.i += 1
            return m

    d5 = {}
    with tempfile.TemporaryDirectory() as td:
        old = os.getcwd()
        os.chdir(td)
        try:
            write_wav_pcm16("speech_TEST.wav", mix_pcm[:L3])
            open("dummy.tflite", "wb").write(b"0" * 1024)
            saved = (ref_tfl.TFLiteBeamformer, ref_fp_inf.TFLiteBeamformer, ref_fp_inf.config.RESULTS_DIR)
            ref_tfl.TFLiteBeamformer = ReplayBeamformer
            ref_fp_inf.TFLiteBeamformer = ReplayBeamformer
            ref_fp_inf.config.RESULTS_DIR = td
            for force64, tag in ((False, "f32read"), (True, "f64read")):
                sf.force_float64 = force64
                sf.written.clear()
                with contextlib.redirect_stdout(io.StringIO()):
                    ref_tfl.process_audio_file("speech_TEST.wav", "paf_out.wav", "dummy.tflite")
                    ref_fp_inf.enhance_audio("g5", "speech_TEST.wav", "dummy.tflite")
                d5["process_audio_file_out_" + tag] = sf.written["paf_out.wav"]
                d5["enhance_audio_out_" + tag] = sf.written["g5_enhanced.wav"]
            sf.force_float64 = False
            ref_tfl.TFLiteBeamformer, ref_fp_inf.TFLiteBeamformer, ref_fp_inf.config.RESULTS_DIR = saved
        finally:
            os.chdir(old)
    # hybrid hard-null with an interference covariance that is exactly zero in a bin above the 200 Hz bypass: the
    # reference divides by zero there (v_int / (v_int[0] / (|v_int[0]| + 1e-10))) and the bin comes out NaN
    rng5 = np.random.default_rng(20261020)
    Yn = ((rng5.standard_normal((2, 513, 16)) + 1j * rng5.standard_normal((2, 513, 16))) * 0.01)
    Yn = Yn.astype(np.complex64).astype(np.complex128)
    mn = rng5.random((513, 16)).astype(np.float32).astype(np.float64)
    mn[40, :] = 1.0
    # (numpy 2.3.5 here: np.linalg.cond of the NaN constraint matrix raises LinAlgError("SVD did not converge"), which
    # hybrid_hard_null_bf does not catch - the reference's answer to such an input is that exception)
    try:
        with np.errstate(all="ignore"):
            d5["hn_nan_out"] = ref_fp_inf.hybrid_hard_null_bf(Yn, mn, f_bins)
        d5["hn_nan_raised"] = np.array("")
    except Exception as ex:                      # noqa: BLE001 - record what the reference does
        d5["hn_nan_raised"] = np.array(f"{type(ex).__name__}: {ex}")
    mn2 = mn.copy()
    mn2[40, :] = rng5.random(16).astype(np.float32).astype(np.float64)
    d5["hn_ok_out"] = ref_fp_inf.hybrid_hard_null_bf(Yn, mn2, f_bins)    # same input without the empty bin
    d5["hn_ok_mask"] = mn2.astype(np.float32)
    d5["hn_nan_Y"] = Yn.astype(np.complex64)
    d5["hn_nan_mask"] = mn.astype(np.float32)
    np.savez_compressed(os.path.join(OUT, "ref_chunk_drivers.npz"), **d5)

    gen_mixer()

    for fn in sorted(os.listdir(OUT)):
        print(fn, os.path.getsize(os.path.join(OUT, fn)))


if __name__ == "__main__":
    main()

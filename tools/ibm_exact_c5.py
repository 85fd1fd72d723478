"""The fused IBM (float32 decision + float64 fix-up of near ties) against the all-float64 decision of EVERY bin, over a
full BASELINE config-5 job: 65 536 synthetic 4 s mixtures = 8.4e9 TF bins.  Inputs are generated
and mixed on the device exactly as tools/sweep_c5.py does.  Writes gpurun_out/ibm_exact_c5.json.

  python tools/ibm_exact_c5.py [n_total]"""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import avzoom  # noqa: E402
from avzoom import ops, pipeline, synth  # noqa: E402
from sweep_c5 import speech_like_batch, FS, DUR_S, N_SRC, BATCH  # noqa: E402


def main():
    n_total = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
    dev = torch.device("cuda", 0)
    cfg = avzoom.PRESETS["baseline_oracle"]
    L = int(DUR_S * FS)
    angles = (synth.TARGET_ANGLE,) + synth.INTERFERER_ANGLES[:N_SRC - 1]
    delays = [synth.far_field_delays(a, 0.04, 343.0) for a in angles]
    gen = torch.Generator(device=dev)
    eng = pipeline.OracleMvdr(cfg, BATCH, L, dev)
    T, F = eng.T, eng.F
    mism, t_exact, t_fused = 0, 0.0, 0.0
    valid = torch.zeros((9,), dtype=torch.int32, device=dev)
    valid[:8] = -1
    valid[8] = 1                                        # bin 256 is bit 0 of word 8; the other 31 bits are unused
    t0 = time.perf_counter()
    for b0 in range(0, n_total, BATCH):
        gen.manual_seed(1_000_003 * 5 + b0)
        src = speech_like_batch(gen, BATCH, N_SRC, L, dev)
        mix, tgt, itf = ops.far_field_mix(src, delays, FS)
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        e0.record()
        eng.pass_a(mix, tgt, itf)
        e1.record()
        exact = ops.ibm_exact_bits(tgt, itf, cfg)
        e2.record()
        torch.cuda.synchronize()
        t_fused += e0.elapsed_time(e1)
        t_exact += e1.elapsed_time(e2)
        diff = (eng.bits ^ exact) & valid
        mism += int((diff != 0).sum().item()) if bool((diff != 0).any()) else 0
    res = {"config": "BASELINE config 5: %d synthetic 4 s mixtures (1 target + 3 interferers), n_fft 512 hop 128" % n_total,
           "utterances": n_total, "frames_per_utterance": T, "bins_checked": n_total * T * F,
           "words_with_a_mismatch": mism, "tolerance": "compiled into the library (4e-6; AVZ_IBM_TOL only in -DAVZ_EXPERIMENT builds)",
           "fused_pass_a_ms_total": t_fused, "all_float64_ibm_ms_total": t_exact, "wall_s": time.perf_counter() - t0,
           "what": "k512_ibm (float32) + k512_ibm_fixup (float64 on near ties) vs avz_ibm_exact_f32 (every bin float64)"}
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(res, open("gpurun_out/ibm_exact_c5.json", "w"), indent=1)
    print(json.dumps(res))


if __name__ == "__main__":
    main()

"""BASELINE config 5: 65536 synthetic 4 s mixtures sharded by utterance over the ranks of one box, generated and
mixed ON THE DEVICE (avz_farfield_mix_f32), enhanced (oracle IBM mask-MVDR), scored, and the (65536, 4) score table
all-gathered with NCCL.  Strong scaling: the total is fixed.

  python tools/sweep_c5.py [n_total]                                   # 1 GPU
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/sweep_c5.py

Prints one JSON line on rank 0.  Device time per stage from CUDA events, max over ranks.
"""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.nn.functional as F

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import avzoom  # noqa: E402
from avzoom import ops, parallel, pipeline, synth  # noqa: E402

FS, DUR_S, N_SRC, BATCH = 16000, 4.0, 4, 1024


def speech_like_batch(gen, B, S, L, dev):
    """White noise under a slow random on/off envelope (syllable-rate, ~10 Hz): time-sparse sources so that the IBM
    is non-trivial.  Data generation only - plain torch ops."""
    x = torch.randn((B, S, L), device=dev, generator=gen)
    coarse = torch.rand((B, S, L // 1600 + 2), device=dev, generator=gen)
    env = F.interpolate((coarse > 0.45).float() * coarse, size=L, mode="linear", align_corners=False)
    return x * env


def main():
    n_total = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    cfg = avzoom.PRESETS["baseline_oracle"]
    L = int(DUR_S * FS)
    lo, hi = parallel.shard_range(n_total, rank, world)
    n_local = hi - lo
    angles = (synth.TARGET_ANGLE,) + synth.INTERFERER_ANGLES[:N_SRC - 1]
    delays = [synth.far_field_delays(a, 0.04, 343.0) for a in angles]
    eng = {}
    scores = torch.empty((n_local, 4), dtype=torch.float32, device=dev)
    gen = torch.Generator(device=dev)
    ev = {k: [] for k in ("generate", "mix", "enhance", "score")}

    def stage(name, fn):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        r = fn()
        e1.record()
        ev[name].append((e0, e1))
        return r

    def run_all():
        for k in ev:
            ev[k].clear()
        for b0 in range(0, n_local, BATCH):
            nb = min(BATCH, n_local - b0)
            gen.manual_seed(1_000_003 * 5 + lo + b0)          # seeded by the first utterance index of the batch
            src = stage("generate", lambda: speech_like_batch(gen, nb, N_SRC, L, dev))
            mix, tgt, itf = stage("mix", lambda: ops.far_field_mix(src, delays, FS))
            if nb not in eng:
                eng[nb] = pipeline.OracleMvdr(cfg, nb, L, dev)
            out = stage("enhance", lambda: eng[nb].run(mix, tgt, itf))
            stage("score", lambda: scores[b0:b0 + nb].copy_(ops.sir_scores(out, tgt, itf)))
        return parallel.gather_scores(scores, n_total)

    run_all()                                                   # warm-up (tables, allocator, NCCL)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    all_scores = run_all()
    t1.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ms = {k: sum(a.elapsed_time(b) for a, b in v) for k, v in ev.items()}
    ms["total"] = t0.elapsed_time(t1)
    if world > 1:
        t = torch.tensor([ms[k] for k in sorted(ms)], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = dict(zip(sorted(ms), t.tolist()))
    if rank == 0:
        audio_s = n_total * DUR_S
        sc = all_scores.float().cpu().numpy()
        print(json.dumps({
            "workload": f"BASELINE config 5: {n_total} synthetic {DUR_S:g} s mixtures (1 target + {N_SRC - 1} interferers), "
                        "generated + mixed on the device, oracle IBM mask-MVDR, scores all-gathered", "n_gpus": world,
            "scaling": "strong", "utterances_per_rank": n_local, "ms_max_over_ranks": {k: round(v, 2) for k, v in ms.items()},
            "audio_s_per_s_enhance_only": round(audio_s / (ms["enhance"] * 1e-3), 0),
            "audio_s_per_s_whole_job": round(audio_s / (ms["total"] * 1e-3), 0),
            "scores_shape": list(sc.shape), "osir_mean_dB": round(float(np.nanmean(sc[:, 1])), 2),
            "finite_scores": int(np.isfinite(sc).all(axis=1).sum())}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

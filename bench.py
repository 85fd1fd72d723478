#!/usr/bin/env python
"""Benchmark of the mask-driven MVDR hot path (BASELINE.json metric: audio-seconds per second).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE config 2): per GPU 1024 synthetic far-field 2-mic mixtures of 4 s at 16 kHz, 1 target +
3 interferers, oracle-IBM mask-MVDR at n_fft 512 / hop 128.  A "step" is one pass of the fused path
(IBM + covariance -> weights -> beamform + post-filter + iSTFT -> peak normalisation) over that batch.
Utterances are independent, so ranks shard them with no data-path collective (weak scaling); NCCL only
all-gathers the per-utterance scores in the end-to-end leg.

`value` is measured with the inputs resident in HBM; `e2e` is the same metric through the public host-buffer
API (pinned host memory -> device every step, enhanced waveforms + scores back to the host).
`--impl reference` times the reference's CPU algorithm (oracle port, float64, all host cores) on a bounded
sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

FS = 16000
DUR_S = 4.0
N_INTERF = 3
UTT_PER_GPU = 1024
CONFIG_ID = 2
ALGO_BYTES_PER_SAMPLE = 20.0        # read mix 2L + tgt L + int L, write out L, float32
METRIC = "audio-sec/sec mask-MVDR (2-mic,16kHz)"
UNIT = "audio-s/s"


# ------------------------------------------------------------------------------------------ CPU baseline
def _oracle_one(seed_args):
    import oracle as O
    from avzoom import synth
    mix, tgt, itf = synth.make_mixture(*seed_args)
    t0 = time.perf_counter()
    O.oracle_mask_mvdr(mix, tgt, itf, O.PRESETS["baseline_oracle"])
    return time.perf_counter() - t0


def cpu_baseline(n_utt: int | None = None):
    """The reference's algorithm (oracle port of rt_av_zoom/core/oracle_debug.py:42-94, float64, per-bin python
    loops as written) over a bounded sample of the workload, one process per host core."""
    import multiprocessing as mp
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    os.environ.setdefault("MKL_NUM_THREADS", "1")
    os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")
    cores = os.cpu_count() or 1
    if n_utt is None:
        n_utt = UTT_PER_GPU                              # the whole config-2 batch: ~25-30 CPU-s of algorithm time
    L = int(DUR_S * FS)
    args = [(1_000_003 * CONFIG_ID + u, L, N_INTERF) for u in range(n_utt)]
    with mp.get_context("fork").Pool(cores) as pool:
        pool.map(_oracle_one, args[:cores])            # warm the workers (imports, FFT plans)
        t0 = time.perf_counter()
        per = pool.map(_oracle_one, args, chunksize=max(1, n_utt // (cores * 4)))
        wall = time.perf_counter() - t0
    # wall includes generating the inputs in the workers; the algorithm's own time is `per`
    algo_cpu_s = float(sum(per))
    eff_wall = algo_cpu_s / cores
    return {
        "value": n_utt * DUR_S / eff_wall,
        "unit": UNIT,
        "cores": cores,
        "kind": "port",
        "sample": f"{n_utt} utterances x {DUR_S:g} s of config 2, float64 numpy/scipy oracle, {cores} processes x 1 thread; "
                  f"{algo_cpu_s:.1f} CPU-s of algorithm time (single core: {n_utt * DUR_S / algo_cpu_s:.1f} audio-s/s)",
        "wall_s_incl_input_generation": wall,
    }


def run_reference(args):
    """Reference arm: the oracle port (the reference's own algorithm, float64 numpy/scipy) on all host cores.  One
    worker pool for the whole run; each of the K steps is a bounded sample of the config-2 workload, sized so that
    W + K steps stay within ~3 batches (3072 utterances, ~80 CPU-s) in total."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp
    for v in ("OMP_NUM_THREADS", "MKL_NUM_THREADS", "OPENBLAS_NUM_THREADS"):
        os.environ.setdefault(v, "1")
    t_all = time.perf_counter()
    cores = os.cpu_count() or 1
    K, W = max(1, args.steps), max(0, args.warmup)
    n_utt = max(cores, min(UTT_PER_GPU, 3 * UTT_PER_GPU // (K + W)))
    L = int(DUR_S * FS)
    vals, cpu_s = [], 0.0
    with mp.get_context("fork").Pool(cores) as pool:
        nxt = 0
        for step in range(W + K):
            sample = [(1_000_003 * CONFIG_ID + (nxt + u) % UTT_PER_GPU, L, N_INTERF) for u in range(n_utt)]
            nxt += n_utt
            per = pool.map(_oracle_one, sample, chunksize=max(1, n_utt // (cores * 4)))
            if step >= W:
                algo = float(sum(per))
                cpu_s += algo
                vals.append(n_utt * DUR_S / (algo / cores))
    v = statistics.median(vals)
    base = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{K} steps x {n_utt} utterances x {DUR_S:g} s of config 2 (utterance index wraps at {UTT_PER_GPU}), float64 "
                      f"numpy/scipy oracle, {cores} processes x 1 thread; {cpu_s:.1f} CPU-s of algorithm time in the timed steps; "
                      "value = median over steps of utterance-seconds / (summed per-utterance algorithm time / cores)"}
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": K,
        "warmup": W, "ms_per_step": 1e3 * cpu_s / cores / K, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "BASELINE config 2 (bounded sample per step): 4 s 2-ch far-field mixtures, 1 target + 3 "
                               "interferers, oracle IBM mask-MVDR, n_fft 512 hop 128", "utterances_per_step": n_utt,
                   "utterance_s": DUR_S},
        "cpu_baseline": base,
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "total_s": time.perf_counter() - t_all,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.rows = []
        self.proc = None
        self.gpu_index = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                                          "-i", str(self.gpu_index), "-lms", "20"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.rows.append((time.perf_counter(), ln.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, smax, reasons = [], None, set()
        for ts, ln in self.rows:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 9:
                continue
            try:
                clk, mx = float(parts[1]), float(parts[2])
            except ValueError:
                continue
            smax = mx
            if t0 <= ts <= t1 + 0.03:        # a sample describes the ~20 ms before it was printed
                sm.append(clk)
                for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), parts[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons),
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------ ours
def load_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def run_ours(args):
    import torch
    import torch.distributed as dist
    import avzoom
    from avzoom import synth, pipeline

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    numa_cpus = None
    if world > 1:
        # host buffers of the e2e leg on the GPU's own NUMA node (8 ranks x 1.3 GB per step otherwise cross sockets)
        from avzoom.parallel import bind_to_gpu_numa
        visible = os.environ.get("CUDA_VISIBLE_DEVICES")
        phys = int(visible.split(",")[local_rank]) if visible and visible.split(",")[local_rank].isdigit() else local_rank
        numa_cpus = bind_to_gpu_numa(phys)
        dist.init_process_group("nccl", device_id=dev)
    cfg = avzoom.PRESETS["baseline_oracle"]
    L = int(DUR_S * FS)

    # ---- synthetic inputs: this rank's shard of the utterance index space
    cores = os.cpu_count() or 1
    workers = max(1, min(32, cores // max(1, world)))
    if numa_cpus:
        workers = max(1, min(workers, len(numa_cpus)))
    distinct = UTT_PER_GPU if workers >= 8 else 256
    t_gen = time.perf_counter()
    mix_h, tgt_h, itf_h = synth.make_batch(CONFIG_ID, distinct, DUR_S, N_INTERF, start=rank * UTT_PER_GPU, workers=workers)
    reps = UTT_PER_GPU // distinct
    if reps > 1:
        mix_h, tgt_h, itf_h = np.tile(mix_h, (reps, 1, 1)), np.tile(tgt_h, (reps, 1)), np.tile(itf_h, (reps, 1))
    t_gen = time.perf_counter() - t_gen
    mix_p = torch.from_numpy(mix_h).pin_memory()
    tgt_p = torch.from_numpy(tgt_h).pin_memory()
    itf_p = torch.from_numpy(itf_h).pin_memory()
    mix, tgt, itf = mix_p.to(dev), tgt_p.to(dev), itf_p.to(dev)
    B = UTT_PER_GPU
    audio_s_per_step = B * DUR_S * world

    enh = pipeline.OracleMvdr(cfg, B, L, dev)      # pre-allocated buffers, no per-step allocation
    # steady-state serving loop: consecutive steps alternate between two engines on two CUDA streams (every step is
    # still one full pass of all seven kernels over the whole batch; the single-stream time is reported next to it)
    DEPTH = 2
    loop = pipeline.StreamedOracleMvdr(cfg, B, L, dev, depth=DEPTH)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- device-resident throughput
    for _ in range(args.warmup):
        enh.run(mix, tgt, itf)
        loop.submit(mix, tgt, itf)
    loop.join()
    barrier()
    # one stream, one engine: the latency of a single step
    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n_single = max(3, min(args.steps, 20))
    s0.record()
    for _ in range(n_single):
        enh.run(mix, tgt, itf)
    s1.record()
    barrier()
    ms_single = max_over_ranks(s0.elapsed_time(s1)) / n_single
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.25)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_wall0 = time.perf_counter()
    e0.record()
    for _ in range(args.steps):
        loop.submit(mix, tgt, itf)
    loop.join()
    e1.record()
    barrier()
    t_wall1 = time.perf_counter()
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    clocks = sampler.stop(t_wall0, t_wall1) if rank == 0 else None
    ms_per_step = ms_total / args.steps
    value = audio_s_per_step / (ms_per_step * 1e-3)

    # ---- per-kernel timing for the roofline (CUDA events on the launching stream, same resident inputs)
    ktimes = enh.time_each_kernel(mix, tgt, itf, iters=max(3, min(args.steps, 10)))
    peak_gbs, peak_src = load_peak()
    samples = B * L
    # Algorithmic (compulsory) bytes each kernel must move per launch: k512_ibm reads tgt L + int L,
    # k512_cov reads mix 2L, k512_apply writes out L (its input is the spectrum pass A kept: not compulsory traffic;
    # the recomputing variant would read mix 2L), k_peak_normalise reads and writes out L.  The path as a whole:
    # 20 B/sample (ALGO_BYTES_PER_SAMPLE).
    algo = {"k512_ibm": 8.0 * samples, "k512_cov": 8.0 * samples, "k512_apply": 4.0 * samples,
            "k_peak_normalise": 8.0 * samples}
    dom = max(algo, key=lambda k: ktimes.get(k, 0.0))
    achieved = algo[dom] / (ktimes[dom] * 1e-3) / 1e9
    traffic, traffic_src = None, None
    issue = None
    try:  # DRAM bytes per launch of that kernel from the committed ncu --set full capture of the same shape
        tj = json.load(open(os.path.join(ROOT, "profiles", "r1_ncu_traffic.json")))
        traffic = tj["per_launch"][dom]["dram_read_bytes"] + tj["per_launch"][dom]["dram_write_bytes"]
        traffic_src = "profiles/r1_ncu_traffic.json (" + tj["source"] + ")"
        # what actually bounds the transform kernels: warp-instruction issue slots (4 schedulers per SM, one
        # instruction per clock each).  Instruction counts from the same ncu capture, times live.
        sms = torch.cuda.get_device_properties(dev).multi_processor_count
        clk_hz = 1e6 * (((clocks or {}).get("sm_mhz")) or 1965.0)
        slots_per_s = sms * 4 * clk_hz
        issue = {"unit": "fraction of warp-instruction issue slots", "peak_slots_per_s": slots_per_s, "kernels": {
            k: (tj["per_launch"][k]["warp_instructions"] * B / 1024.0) / (ktimes[k] * 1e-3) / slots_per_s
            for k in ("k512_ibm", "k512_cov", "k512_apply") if ktimes.get(k)},
            "note": "instruction counts per launch from profiles/r1_ncu_traffic.json (ncu smsp__inst_executed.sum at "
                    "B = 1024), divided by the CUDA-event kernel time measured in this run"}
    except Exception:
        pass
    roofline = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak_gbs, "unit": "GB/s",
                "frac": achieved / peak_gbs, "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": algo[dom], "kernel_ms": ktimes,
                "note": "fp32-issue-bound path (DESIGN.md 3.1): ~2540 warp-instructions per frame set the time, not HBM",
                "path_achieved_GBps": (value / world) * FS * ALGO_BYTES_PER_SAMPLE / 1e9,
                "path_frac_of_hbm_roofline": (value / world) * FS * ALGO_BYTES_PER_SAMPLE / (peak_gbs * 1e9),
                "issue_slots": issue}

    # ---- end to end through the public host-buffer API
    e2e_steps = max(2, min(args.steps, 5))
    host = pipeline.HostPipeline(enh, world)
    host.run(mix_p, tgt_p, itf_p)                       # warm-up (allocates pinned result buffers)
    barrier()
    t0 = time.perf_counter()
    g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g0.record()
    for _ in range(e2e_steps):
        out_h, scores_h = host.run(mix_p, tgt_p, itf_p)
    g1.record()
    barrier()
    e2e_ms = max_over_ranks(g0.elapsed_time(g1)) / e2e_steps
    e2e = {"value": audio_s_per_step / (e2e_ms * 1e-3), "unit": UNIT, "ms_per_step": e2e_ms,
           "h2d_bytes_per_step": int(mix_p.numel() + tgt_p.numel() + itf_p.numel()) * 4,
           "d2h_bytes_per_step": int(out_h.numel() + scores_h.numel()) * 4,
           "api": "avzoom.pipeline.HostPipeline.run(pinned mix, tgt, itf) -> (enhanced waveforms, all-gathered scores)"}
    sir_mean = float(scores_h[:, 1].mean())
    # same leg with the reference's on-disk sample format (PCM16 WAV, oracle_debug.py:35-39,96) on the wire: int16 in
    # both directions, converted on the device.  Informational: the headline e2e above moves float32.
    del host
    to_pcm = lambda a: torch.from_numpy(np.clip(np.rint(a * 32767.0), -32768, 32767).astype(np.int16)).pin_memory()
    mix_w, tgt_w, itf_w = to_pcm(mix_h), to_pcm(tgt_h), to_pcm(itf_h)
    host_w = pipeline.HostPipeline(enh, world, wire="pcm16")
    host_w.run(mix_w, tgt_w, itf_w)
    barrier()
    g0.record()
    for _ in range(e2e_steps):
        out_w, scores_w = host_w.run(mix_w, tgt_w, itf_w)
    g1.record()
    barrier()
    w_ms = max_over_ranks(g0.elapsed_time(g1)) / e2e_steps
    e2e["pcm16_wire"] = {"value": audio_s_per_step / (w_ms * 1e-3), "unit": UNIT, "ms_per_step": w_ms,
                         "h2d_bytes_per_step": int(mix_w.numel() + tgt_w.numel() + itf_w.numel()) * 2,
                         "d2h_bytes_per_step": int(out_w.numel()) * 2 + int(scores_w.numel()) * 4,
                         "output_sir_mean": float(scores_w[:, 1].mean())}
    del host_w, mix_w, tgt_w, itf_w
    sir_in = float(avzoom.sir_scores(mix[:, 0, :].contiguous(), tgt, itf)[:, 1].mean())

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline()

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": "BASELINE config 2: 1024 synthetic 4 s 2-ch far-field mixtures per GPU, 1 target + 3 "
                                   "interferers, oracle IBM mask-MVDR, n_fft 512 hop 128",
                       "utterances_per_gpu": B, "distinct_utterances_per_gpu": distinct, "samples_per_utterance": L,
                       "l2_policy": "inputs (1.05 GB per GPU) larger than the 126 MB L2; no flush needed",
                       "step_schedule": f"steps alternate between {DEPTH} engines on {DEPTH} CUDA streams (steady-state "
                                        "serving loop); each step is one full pass of the 7 kernels over the whole batch",
                       "ms_per_step_single_stream": ms_single,
                       "input_generation_s": round(t_gen, 2),
                       "host_cpus_bound_to_gpu_numa_node": len(numa_cpus) if numa_cpus else None},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e,
            "gpu_launches": enh.launches_per_step * args.steps, "clocks": clocks,
            "dSIR_dB": {"output_sir_mean": sir_mean, "mic1_sir_mean": sir_in, "improvement": sir_mean - sir_in},
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        if args.gpus > 1 and "RANK" not in os.environ:
            cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
                   "--master-addr", "127.0.0.1", "--master-port", "29511", os.path.abspath(__file__), "--gpus", str(args.gpus),
                   "--steps", str(args.steps), "--warmup", str(args.warmup)]
            sys.exit(subprocess.call(cmd))
        run_ours(args)


if __name__ == "__main__":
    main()

"""The non-headline BASELINE configurations as functions, so that bench.py can append a compact block for each to its
JSON line and the driver witnesses them.  Device time from CUDA events throughout.

  config1  one 5 s mixture (the reference's own CPU-runnable case): step latency, eager and as a CUDA graph
  config3  256 x 4 s: features -> random-init U-Net -> learned-mask MVDR, the hot-path stages and the U-Net apart;
           plus the chunk-driver form (2 s windows in place, chunk OLA kernel) over 256 recordings
  config4  4096 concurrent streams: per-hop latency p50 / p99
  config5  65536 mixtures generated + mixed on the device, enhanced, scored, score table all-gathered (all ranks)
"""
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import avzoom  # noqa: E402
from avzoom import ops, parallel, pipeline, stream, synth  # noqa: E402

FS = 16000


def _timed(fn, iters):
    r = None
    for _ in range(3):        # warm-up: kernels, and the two output buffers the loop below alternates between
        r = fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        r = fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters, r


def config1(dev, iters: int = 200):
    cfg = avzoom.PRESETS["baseline_oracle"]
    mix, tgt, itf = synth.make_batch(1, 1, 5.0, 2)
    L = mix.shape[-1]
    mix_d, tgt_d, itf_d = (torch.from_numpy(a).to(dev) for a in (mix, tgt, itf))
    eng = pipeline.OracleMvdr(cfg, 1, L, dev)
    for _ in range(5):
        eng.run(mix_d, tgt_d, itf_d)
    torch.cuda.synchronize()
    ref = eng.out.clone()
    eager_ms, _ = _timed(lambda: eng.run(mix_d, tgt_d, itf_d), iters)
    side = torch.cuda.Stream(dev)
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        eng.run(mix_d, tgt_d, itf_d)
    torch.cuda.current_stream().wait_stream(side)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        eng.run(mix_d, tgt_d, itf_d)
    graph_ms, _ = _timed(g.replay, iters)
    same = bool(torch.equal(eng.out, ref))
    return {"workload": "one 5 s 2-mic mixture, 1 target + 2 interferers, oracle IBM mask-MVDR, n_fft 512 hop 128",
            "eager_us": round(eager_ms * 1e3, 2), "cuda_graph_us": round(graph_ms * 1e3, 2),
            "graph_output_bit_identical": same, "times_real_time_graph": round(5.0 / (graph_ms * 1e-3))}


class _CheapMask(torch.nn.Module):
    def forward(self, X):
        return torch.sigmoid(0.5 * X[:, 0] + 3.0 + 0.3 * torch.cos(X[:, 1]))


def config3(dev, B: int = 256, unet_iters: int = 2):
    from avzoom.core import models, chunked
    cfg = avzoom.PRESETS["baseline_learned"]
    mix8, _, _ = synth.make_batch(3, 8, 4.0, 3)
    mix = torch.from_numpy(mix8).to(dev).repeat((B + 7) // 8, 1, 1)[:B].contiguous()
    torch.manual_seed(0)
    net = models.FreqPreservingUNet().eval().to(dev)
    t_feat, X = _timed(lambda: avzoom.wave_features(mix, cfg.n_fft, cfg.hop), 10)
    sub = 16

    def unet():
        with torch.no_grad():
            return torch.cat([net(X[i:i + sub]) for i in range(0, B, sub)]).float().contiguous()

    t_net, mask = _timed(unet, unet_iters)
    # the same network in bf16 autocast + channels-last (the mask model is outside the path's scope - SURVEY 8-A lists it
    # as the consumer of the features; reported so that the split of a config-3 step is visible)
    net_cl = net.to(memory_format=torch.channels_last)

    def unet_bf16():
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
            return torch.cat([net_cl(X[i:i + sub].contiguous(memory_format=torch.channels_last)) for i in range(0, B, sub)]).float().contiguous()

    try:
        t_net16, mask16 = _timed(unet_bf16, unet_iters)
        mask_err = float((mask16 - mask).abs().max())
        del mask16
    except Exception as ex:  # noqa: BLE001
        t_net16, mask_err = None, repr(ex)
    t_mvdr, out = _timed(lambda: avzoom.learned_mask_mvdr(mix, mask, cfg), 10)
    del X, mask
    # the reference's own deployment form: 2 s windows at 50 % overlap, n_fft 1024 / hop 512, count-averaged chunk OLA
    enh = chunked.ChunkedEnhancer(avzoom.PRESETS["full_audio"])
    cheap = _CheapMask()
    for _ in range(3):
        enh(mix, cheap, timing=True)
    ck = enh.last_ms
    audio_s = B * 4.0
    return {"workload": f"{B} x 4 s: log-mag + IPD features -> FreqPreservingUNet (random init, eval, torch fp32) -> "
                        "learned-mask MVDR (sigma 1e-5, post-filter max(M, 0.05)), n_fft 512 hop 128",
            "features_ms": round(t_feat, 4), "mask_mvdr_ms": round(t_mvdr, 4), "unet_ms": round(t_net, 2),
            "unet_bf16_channels_last_ms": None if t_net16 is None else round(t_net16, 2), "unet_bf16_max_mask_diff": mask_err,
            "hot_path_audio_s_per_s": round(audio_s / ((t_feat + t_mvdr) * 1e-3)),
            "with_unet_audio_s_per_s": round(audio_s / ((t_feat + t_net + t_mvdr) * 1e-3)),
            "finite": bool(torch.isfinite(out).all()),
            "chunk_driver": {"what": f"{B} recordings x 4 s as 2 s windows read in place (n_fft 1024 hop 512), pointwise "
                                     "stand-in mask model, count-averaged chunk OLA kernel",
                             "features_ms": round(ck["features"], 4), "mvdr_and_chunk_ola_ms": round(ck["mvdr_and_chunk_ola"], 4),
                             "hot_path_audio_s_per_s": round(audio_s / ((ck["features"] + ck["mvdr_and_chunk_ola"]) * 1e-3))}}


def config4(dev, S: int = 4096, hops: int = 500):
    eng = stream.MvdrStream(S, avzoom.PRESETS["baseline_oracle"], lam=0.95, device=dev)
    x = torch.randn((S, 2, 128), device=dev) * 0.1
    m = torch.rand((S, 257), device=dev)
    for _ in range(20):
        eng.step(x, m)
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(hops)]
    for a, b in ev:
        a.record()
        eng.step(x, m)
        b.record()
    torch.cuda.synchronize()
    t = np.array([a.elapsed_time(b) for a, b in ev]) * 1e3
    state = int(eng.lib.avz_stream_state_bytes(1))
    algo = 2 * state + 1024 + 512 + 257 * 4
    return {"workload": f"{S} concurrent 2-mic streams, recursive covariance (lambda 0.95), one 128-sample hop per call, "
                        "n_fft 512 (NOT in the reference: parity unpinned)",
            "hop_us_p50": round(float(np.percentile(t, 50)), 2), "hop_us_p99": round(float(np.percentile(t, 99)), 2),
            "stream_hops_per_s": round(S / (t.mean() * 1e-6)), "times_real_time_p99": round(8000.0 / float(np.percentile(t, 99)), 1),
            "state_traffic_GBps": round(algo * S / (t.mean() * 1e-6) / 1e9, 1)}


def _speech_like_batch(gen, B, S, L, dev):
    x = torch.randn((B, S, L), device=dev, generator=gen)
    coarse = torch.rand((B, S, L // 1600 + 2), device=dev, generator=gen)
    env = F.interpolate((coarse > 0.45).float() * coarse, size=L, mode="linear", align_corners=False)
    return x * env


def config5(dev, rank: int, world: int, n_total: int = 65536, batch: int = 1024):
    import torch.distributed as dist
    cfg = avzoom.PRESETS["baseline_oracle"]
    L, n_src = 4 * FS, 4
    lo, hi = parallel.shard_range(n_total, rank, world)
    n_local = hi - lo
    angles = (synth.TARGET_ANGLE,) + synth.INTERFERER_ANGLES[:n_src - 1]
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
        for b0 in range(0, n_local, batch):
            nb = min(batch, n_local - b0)
            gen.manual_seed(1_000_003 * 5 + lo + b0)
            src = stage("generate", lambda: _speech_like_batch(gen, nb, n_src, L, dev))
            mix, tgt, itf = stage("mix", lambda: ops.far_field_mix(src, delays, FS))
            if nb not in eng:
                eng[nb] = pipeline.OracleMvdr(cfg, nb, L, dev)
            out = stage("enhance", lambda: eng[nb].run(mix, tgt, itf))
            stage("score", lambda: scores[b0:b0 + nb].copy_(ops.sir_scores(out, tgt, itf)))
        return parallel.gather_scores(scores, n_total)

    run_all()
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
    sc = all_scores.float().cpu().numpy()
    audio_s = n_total * 4.0
    return {"workload": f"{n_total} synthetic 4 s mixtures (1 target + 3 interferers) sharded by utterance over {world} rank(s), "
                        "generated + mixed on the device, oracle IBM mask-MVDR, (n, 4) score table all-gathered (NCCL)",
            "scaling": "strong", "ms_max_over_ranks": {k: round(v, 1) for k, v in ms.items()},
            "whole_job_audio_s_per_s": round(audio_s / (ms["total"] * 1e-3)),
            "enhance_only_audio_s_per_s": round(audio_s / (ms["enhance"] * 1e-3)),
            "scores_shape": list(sc.shape), "finite_score_rows": int(np.isfinite(sc).all(axis=1).sum()),
            "osir_mean_dB": round(float(np.nanmean(sc[:, 1])), 2)}

"""Batch drivers around the fused kernels: a pre-allocated oracle-mask MVDR engine (no per-step allocation, so a
step is exactly the kernel launches) and a host-buffer pipeline that overlaps PCIe copies with compute.

Utterances are independent (each has its own covariance), so a multi-GPU run shards them by contiguous blocks
of the utterance index and never exchanges waveform data; the only collective is an all-gather of the
per-utterance scores.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Tuple

import torch

from . import _lib
from .config import MvdrConfig
from .ops import _ptr, _stream, num_frames, steering_vectors


class OracleMvdr:
    """Oracle-IBM mask-MVDR (rt_av_zoom/core/oracle_debug.py:42-94) for a fixed batch shape."""

    # k512_ibm, k512_ibm_fixup, k512_cov, k_cov_finalize, k_mvdr_weights, k512_apply, k_peak_normalise
    # (+ two small memsets: the near-tie counter and `peak`)
    launches_per_step = 7

    def __init__(self, cfg: MvdrConfig, B: int, L: int, device, keep_spectrum: bool = True, fused_norm: bool = False,
                 fused: bool = False, fold_weights: bool = False, sparse_spectrum: bool = False,
                 skip_masked_stores: bool = True):
        """`fused`: run pass A, the weights, pass B and the normalisation as ONE persistent kernel whose kept spectrum
        stays in L2 (avz_oracle_fused_f32; n_fft 512, IBM post-filter or none) instead of five separate launches."""
        self.cfg, self.B, self.L, self.device = cfg, B, L, device
        self.lib = _lib.load()
        self.F = cfg.n_freq
        self.T = num_frames(L, cfg.n_fft, cfg.hop)
        self.out_len = (self.T - 1) * cfg.hop
        f32 = dict(dtype=torch.float32, device=device)
        self.bits = torch.empty((B, self.T, (self.F + 31) // 32), dtype=torch.int32, device=device)
        self.R = torch.empty((B, self.F, 4), **f32)
        self.msum = torch.empty((B, self.F), **f32)
        self.w = torch.empty((B, self.F, 2), dtype=torch.complex64, device=device)
        self.out = torch.empty((B, self.out_len), **f32)
        self.peak = torch.zeros((B,), **f32)
        nws = self.lib.avz_ibm_cov_ws_bytes(B, L, cfg.n_fft, cfg.hop)
        if nws < 0:
            _lib.check(-1, "avz_ibm_cov_ws_bytes")
        self.ws = torch.empty((max(int(nws), 4),), dtype=torch.uint8, device=device)
        # pass A may keep the packed mix spectrum so that pass B skips its forward transform (fast path only)
        nspec = self.lib.avz_spec_ws_bytes(B, L, cfg.n_fft, cfg.hop) if keep_spectrum and cfg.n_fft == 512 else 0
        # zeroed once: the store-skipping pass A (below) leaves sectors unwritten that pass B multiplies by a gain of 0
        self.spec = torch.zeros((int(nspec),), dtype=torch.uint8, device=device) if nspec > 0 else None
        # fused_norm: the thread blocks of an utterance run as one cluster, exchange their maxima through distributed
        # shared memory and each rescales the range it has just written while it is still in L2 (one launch less,
        # bit-identical output).  Measured on B200 at config 2 it LOSES: pass B 0.50 -> 0.68 ms against the 0.09 ms of
        # the separate streaming kernel (L2 round trips at the end of every block + cluster co-scheduling).  Off.
        self.fused_norm = fused_norm and cfg.peak_eps is not None and self.spec is not None
        self.launches_per_step = 6 if self.fused_norm else 7
        self.d = steering_vectors(cfg, device)
        self.cc = cfg.to_c()
        self.fused = bool(fused)
        # oracle post-filter (1 - noise mask): pass B zeroes every noise-dominated bin, so pass A can keep only the others
        # (avz_ibm_cov_keep_sparse_f32 / avz_mvdr_apply_kept_sparse_f32): the kept-spectrum traffic of the two passes
        # falls from 5.05 to 1.9 GB per 1024 x 4 s with the same bits out - but the compaction costs ~120 / ~200 more
        # instructions per frame, and both kernels are then issue-bound: 0.64 + 0.69 ms against 0.57 + 0.47 dense
        # (profiles/README.md).  Off by default.
        self.sparse = bool(sparse_spectrum) and self.spec is not None and cfg.post == "one_minus_noise" and not self.fused_norm
        # ... what does pay: the dense layout with the stores of sectors pass B will zero anyway left out
        # (avz_ibm_cov_keep_postmask_f32): about half of pass A's store stream never goes to HBM
        self.skip = (bool(skip_masked_stores) and self.spec is not None and cfg.post == "one_minus_noise"
                     and cfg.n_fft == 512 and not self.sparse)
        # n_fft 512 fast path: finalize + weights folded into pass A's last block per utterance (two launches fewer,
        # bit-identical).  Measured on B200 it does not pay: a single block per utterance is slower at the tail than the
        # two small parallel kernels it replaces (config 2: 1.699 vs 1.696 ms per step; one 5 s utterance as a CUDA
        # graph: 59.5 vs 53.3 us).  Off by default.
        self.fold_weights = bool(fold_weights) and cfg.n_fft == 512 and cfg.hop in (128, 256) and not self.fused
        if self.fold_weights:
            self.launches_per_step -= 2
        if self.fused:
            nfw = self.lib.avz_oracle_fused_ws_bytes(B, L, cfg.n_fft, cfg.hop)
            if nfw <= 0 or cfg.post not in ("one_minus_noise", "none"):
                raise _lib.AvzError("the fused oracle kernel needs n_fft 512 (hop 128 / 256) and an IBM or no post-filter")
            self.spec = None          # the fused kernel keeps its spectrum in a ring inside its own workspace
            self.fws = torch.empty((int(nfw),), dtype=torch.uint8, device=device)
            self.launches_per_step = 3
        _lib.check(self.lib.avz_init(cfg.n_fft), "avz_init")

    # individual stages (each one C-ABI call) ---------------------------------------------------
    def pass_a(self, mix, tgt, itf):
        c = self.cfg
        if self.fold_weights:
            # finalize + 2x2 solve ride on pass A's last block per utterance: R, msum and w come out of this one call
            _lib.check(self.lib.avz_ibm_cov_weights_keep_f32(_ptr(mix), _ptr(tgt), _ptr(itf), self.B, self.L, c.n_fft, c.hop,
                                                             C.byref(self.cc), _ptr(self.d), _ptr(self.bits), _ptr(self.R),
                                                             _ptr(self.msum), _ptr(self.w), _ptr(self.ws), _ptr(self.spec),
                                                             1 if self.sparse else (2 if self.skip else 0), _stream()),
                       "avz_ibm_cov_weights_keep_f32")
            return
        if self.spec is not None and self.sparse:
            _lib.check(self.lib.avz_ibm_cov_keep_sparse_f32(_ptr(mix), _ptr(tgt), _ptr(itf), self.B, self.L, c.n_fft, c.hop,
                                                            float(c.norm_eps), _ptr(self.bits), _ptr(self.R), _ptr(self.msum),
                                                            _ptr(self.ws), _ptr(self.spec), _stream()),
                       "avz_ibm_cov_keep_sparse_f32")
            return
        if self.spec is not None and self.skip:
            _lib.check(self.lib.avz_ibm_cov_keep_postmask_f32(_ptr(mix), _ptr(tgt), _ptr(itf), self.B, self.L, c.n_fft, c.hop,
                                                              float(c.norm_eps), _ptr(self.bits), _ptr(self.R), _ptr(self.msum),
                                                              _ptr(self.ws), _ptr(self.spec), _stream()),
                       "avz_ibm_cov_keep_postmask_f32")
            return
        if self.spec is not None:
            _lib.check(self.lib.avz_ibm_cov_keep_f32(_ptr(mix), _ptr(tgt), _ptr(itf), self.B, self.L, c.n_fft, c.hop,
                                                     float(c.norm_eps), _ptr(self.bits), _ptr(self.R), _ptr(self.msum),
                                                     _ptr(self.ws), _ptr(self.spec), _stream()), "avz_ibm_cov_keep_f32")
            return
        _lib.check(self.lib.avz_ibm_cov_f32(_ptr(mix), _ptr(tgt), _ptr(itf), self.B, self.L, c.n_fft, c.hop,
                                            float(c.norm_eps), _ptr(self.bits), _ptr(self.R), _ptr(self.msum),
                                            _ptr(self.ws), _stream()), "avz_ibm_cov_f32")

    def weights(self):
        if self.fold_weights:
            return                      # done by pass_a
        _lib.check(self.lib.avz_mvdr_weights_f32(_ptr(self.R), _ptr(self.d), self.B, self.F, C.byref(self.cc),
                                                 _ptr(self.w), _stream()), "avz_mvdr_weights_f32")

    def pass_b(self, mix, out=None):
        c = self.cfg
        self.peak.zero_()
        out_t = self.out if out is None else out
        bits = self.bits if c.post == "one_minus_noise" else None
        if self.spec is not None and self.fused_norm:
            # peak normalisation fused into pass B (thread-block cluster per utterance)
            _lib.check(self.lib.avz_mvdr_apply_kept_norm_f32(_ptr(self.spec), _ptr(self.w), _ptr(bits), _ptr(None),
                                                             self.B, self.L, c.n_fft, c.hop, C.byref(self.cc),
                                                             float(c.peak_eps), _ptr(out_t), _ptr(self.peak),
                                                             _stream()), "avz_mvdr_apply_kept_norm_f32")
            return
        if self.spec is not None and self.sparse:
            _lib.check(self.lib.avz_mvdr_apply_kept_sparse_f32(_ptr(self.spec), _ptr(self.w), _ptr(bits), self.B, self.L,
                                                               c.n_fft, c.hop, C.byref(self.cc), _ptr(out_t), _ptr(self.peak),
                                                               _stream()), "avz_mvdr_apply_kept_sparse_f32")
            return
        if self.spec is not None:
            _lib.check(self.lib.avz_mvdr_apply_kept_f32(_ptr(self.spec), _ptr(self.w), _ptr(bits), _ptr(None), self.B,
                                                        self.L, c.n_fft, c.hop, C.byref(self.cc), _ptr(out_t),
                                                        _ptr(self.peak), _stream()), "avz_mvdr_apply_kept_f32")
            return
        _lib.check(self.lib.avz_mvdr_apply_f32(_ptr(mix), _ptr(self.w), _ptr(bits), _ptr(None), self.B, self.L, c.n_fft,
                                               c.hop, C.byref(self.cc), _ptr(out_t), _ptr(self.peak), _stream()),
                   "avz_mvdr_apply_f32")

    def normalise(self, out=None):
        if self.cfg.peak_eps is not None and not (self.spec is not None and self.fused_norm):
            _lib.check(self.lib.avz_peak_normalise_f32(_ptr(self.out if out is None else out), self.B, self.out_len, _ptr(self.peak),
                                                       float(self.cfg.peak_eps), _stream()), "avz_peak_normalise_f32")

    def run(self, mix: torch.Tensor, tgt: torch.Tensor, itf: torch.Tensor, out: torch.Tensor = None) -> torch.Tensor:
        """One step over device-resident inputs mix [B,2,L], tgt [B,L], itf [B,L] -> out [B,(T-1)*hop] (the engine's own
        buffer, or the contiguous float32 tensor `out` of that shape)."""
        if out is not None and (out.shape != self.out.shape or out.dtype != torch.float32 or not out.is_contiguous()):
            raise ValueError("out must be a contiguous float32 tensor of shape %s" % (tuple(self.out.shape),))
        if self.fused:
            c = self.cfg
            dst = self.out if out is None else out
            _lib.check(self.lib.avz_oracle_fused_f32(_ptr(mix), _ptr(tgt), _ptr(itf), self.B, self.L, c.n_fft, c.hop,
                                                     C.byref(self.cc), -1.0 if c.peak_eps is None else float(c.peak_eps),
                                                     _ptr(self.d), _ptr(self.bits), _ptr(self.R), _ptr(self.msum), _ptr(self.w),
                                                     _ptr(dst), _ptr(self.peak), _ptr(self.fws), _stream()),
                       "avz_oracle_fused_f32")
            return dst
        self.pass_a(mix, tgt, itf)
        self.weights()
        self.pass_b(mix, out)
        self.normalise(out)
        return self.out if out is None else out

    KERNELS = ("k512_ibm", "k512_ibm_fixup", "k512_cov", "k_cov_finalize", "k_mvdr_weights", "k512_apply",
               "k_peak_normalise")
    FUSED_KERNELS = ("k512_ibm", "k512_ibm_fixup", "k512_fused")

    def time_each_kernel(self, mix, tgt, itf, iters: int = 5) -> Dict[str, float]:
        """Average device time (ms) of every kernel of a step, from CUDA events the library records on the launching
        stream around each launch (avz_profile_enable / avz_profile_get)."""
        self.run(mix, tgt, itf)
        torch.cuda.synchronize()
        _lib.check(self.lib.avz_profile_enable(1), "avz_profile_enable")
        acc = [0.0] * len(self.KERNELS)
        buf = (C.c_float * len(self.KERNELS))()
        try:
            for _ in range(iters):
                self.run(mix, tgt, itf)
                _lib.check(self.lib.avz_profile_get(buf, len(self.KERNELS)), "avz_profile_get")
                for i in range(len(self.KERNELS)):
                    acc[i] += max(0.0, float(buf[i]))
        finally:
            self.lib.avz_profile_enable(0)
        if self.fused:      # the persistent kernel is timed in pass A's slot
            return {k: acc[i] / iters for i, k in enumerate(self.FUSED_KERNELS)}
        return {k: acc[i] / iters for i, k in enumerate(self.KERNELS)}

    def time_kernels(self, mix, tgt, itf, iters: int = 5) -> Dict[str, float]:
        """Average device time (ms) of each stage, CUDA events on the launching stream."""
        stages = [("pass A (k512_ibm + k512_ibm_fixup + k512_cov + k_cov_finalize)", lambda: self.pass_a(mix, tgt, itf)),
                  ("k_mvdr_weights", self.weights),
                  ("pass B (k512_apply)", lambda: self.pass_b(mix)),
                  ("k_peak_normalise", self.normalise)]
        self.run(mix, tgt, itf)
        torch.cuda.synchronize()
        res = {}
        for name, fn in stages:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(iters):
                fn()
            e1.record()
            torch.cuda.synchronize()
            res[name] = e0.elapsed_time(e1) / iters
        return res


class StreamedOracleMvdr:
    """Steady-state serving loop: consecutive batches go to `depth` engines on their own CUDA streams, so the
    issue-bound first kernel of one batch meets the latency-bound passes of the previous one on the SMs (+4 % batches
    per second at BASELINE config 2 with depth 2; every batch is still one full pass of the seven kernels).

    submit() enqueues a batch and returns the tensor its result will be in; join() makes the caller's stream wait for
    everything submitted.  A result tensor is reused after `depth` further submits."""

    def __init__(self, cfg: MvdrConfig, B: int, L: int, device, depth: int = 2, fused: bool = False):
        if depth < 1:
            raise ValueError("depth must be >= 1")
        self.engines = [OracleMvdr(cfg, B, L, device, fused=fused) for _ in range(depth)]
        self.streams = [torch.cuda.Stream(device) for _ in range(depth)]
        self.launches_per_step = self.engines[0].launches_per_step
        self.n = 0

    def submit(self, mix: torch.Tensor, tgt: torch.Tensor, itf: torch.Tensor) -> torch.Tensor:
        i = self.n % len(self.engines)
        self.n += 1
        s = self.streams[i]
        s.wait_stream(torch.cuda.current_stream())      # the inputs are ready on the caller's stream
        with torch.cuda.stream(s):
            return self.engines[i].run(mix, tgt, itf)

    def join(self) -> None:
        cur = torch.cuda.current_stream()
        for s in self.streams:
            cur.wait_stream(s)


class HostPipeline:
    """End-to-end leg: pinned host waveforms in, enhanced waveforms + scores out, every step.

    The batch is cut into sub-batches; host->device copies, kernels and device->host copies of consecutive
    sub-batches run on three streams so PCIe transfers overlap compute.  Scores (OSINR, OSIR, SDR, SIR per
    utterance) are all-gathered across ranks with NCCL when torch.distributed is initialised.

    submit() enqueues a whole step and returns a ticket without waiting; wait(ticket) blocks until that step's results
    are in pinned host memory.  Steps pipeline: the first host->device copy of step k+1 runs while step k is still
    computing and draining, so in steady state the leg runs at the speed of its slowest resource (the host->device
    link) with no fill/drain bubble per step.  Results live in `host_slots` rotating pinned buffers: a ticket's
    buffers are reused `host_slots` submits later.  run() = submit + wait.

    wire="pcm16": the host buffers are int16 (what the reference's WAV files hold, oracle_debug.py:35-39,96); samples
    cross PCIe as int16 in both directions and are converted on the device (read = /32768 like soundfile, write =
    round(x*32767) like libsndfile), halving the transfer that bounds this leg."""

    def __init__(self, engine: OracleMvdr, world: int = 1, sub_batches: int = 8, wire: str = "f32", host_slots: int = 2):
        if wire not in ("f32", "pcm16"):
            raise ValueError("wire must be 'f32' or 'pcm16'")
        self.e = engine
        self.wire = wire
        self.world = world
        B = engine.B
        self.nsub = sub_batches if B % sub_batches == 0 and B >= sub_batches else 1
        self.sb = B // self.nsub
        dev = engine.device
        self.sub = OracleMvdr(engine.cfg, self.sb, engine.L, dev, fused=engine.fused)
        f32 = dict(dtype=torch.float32, device=dev)
        self.d_mix = [torch.empty((self.sb, 2, engine.L), **f32) for _ in range(2)]
        self.d_tgt = [torch.empty((self.sb, engine.L), **f32) for _ in range(2)]
        self.d_itf = [torch.empty((self.sb, engine.L), **f32) for _ in range(2)]
        self.d_out = [torch.empty((self.sb, engine.out_len), **f32) for _ in range(2)]
        self.host_slots = max(1, int(host_slots))
        self.scores = [torch.empty((B, 4), **f32) for _ in range(self.host_slots)]
        self.scores_all = [torch.empty((B * world, 4), **f32) for _ in range(self.host_slots)]
        if wire == "pcm16":
            i16 = dict(dtype=torch.int16, device=dev)
            self.w_mix = [torch.empty((self.sb, 2, engine.L), **i16) for _ in range(2)]
            self.w_tgt = [torch.empty((self.sb, engine.L), **i16) for _ in range(2)]
            self.w_itf = [torch.empty((self.sb, engine.L), **i16) for _ in range(2)]
            self.w_out = [torch.empty((self.sb, engine.out_len), **i16) for _ in range(2)]
        odt = torch.int16 if wire == "pcm16" else torch.float32
        self.h_out = [torch.empty((B, engine.out_len), dtype=odt).pin_memory() for _ in range(self.host_slots)]
        self.h_scores = [torch.empty((B * world, 4), dtype=torch.float32).pin_memory() for _ in range(self.host_slots)]
        self.s_in, self.s_cmp, self.s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev), torch.cuda.Stream(dev)
        # events that order the reuse of the two device slots; they persist across steps so that steps pipeline
        self.ev_in = [None, None]
        self.ev_cmp = [None, None]
        self.ev_out = [None, None]
        self.done = [None] * self.host_slots
        self.n_submitted = 0

    def submit(self, mix_h: torch.Tensor, tgt_h: torch.Tensor, itf_h: torch.Tensor) -> int:
        lib = _lib.load()
        pcm = self.wire == "pcm16"
        want = torch.int16 if pcm else torch.float32
        if mix_h.dtype != want or tgt_h.dtype != want or itf_h.dtype != want:
            raise _lib.AvzError(f"HostPipeline(wire={self.wire!r}) needs {want} host buffers")
        ticket = self.n_submitted
        hs = ticket % self.host_slots
        self.n_submitted += 1
        i_mix, i_tgt, i_itf = (self.w_mix, self.w_tgt, self.w_itf) if pcm else (self.d_mix, self.d_tgt, self.d_itf)
        cur = torch.cuda.current_stream()
        for s in (self.s_in, self.s_cmp, self.s_out):
            s.wait_stream(cur)
        ev_in, ev_cmp, ev_out = self.ev_in, self.ev_cmp, self.ev_out
        scores, h_out = self.scores[hs], self.h_out[hs]
        for c in range(self.nsub):
            slot = c & 1
            lo, hi = c * self.sb, (c + 1) * self.sb
            with torch.cuda.stream(self.s_in):
                if ev_cmp[slot] is not None:
                    self.s_in.wait_event(ev_cmp[slot])       # previous user of this slot's inputs has finished
                i_mix[slot].copy_(mix_h[lo:hi], non_blocking=True)
                i_tgt[slot].copy_(tgt_h[lo:hi], non_blocking=True)
                i_itf[slot].copy_(itf_h[lo:hi], non_blocking=True)
                ev_in[slot] = self.s_in.record_event()
            with torch.cuda.stream(self.s_cmp):
                self.s_cmp.wait_event(ev_in[slot])
                if ev_out[slot] is not None:
                    self.s_cmp.wait_event(ev_out[slot])      # previous output of this slot has left the device
                if pcm:
                    for w, d in ((i_mix[slot], self.d_mix[slot]), (i_tgt[slot], self.d_tgt[slot]), (i_itf[slot], self.d_itf[slot])):
                        _lib.check(lib.avz_pcm16_to_f32(_ptr(w), w.numel(), _ptr(d), _stream()), "avz_pcm16_to_f32")
                # the engine writes straight into the device slot the device->host copy reads (no staging copy)
                out = self.sub.run(self.d_mix[slot], self.d_tgt[slot], self.d_itf[slot], out=self.d_out[slot])
                _lib.check(lib.avz_sir_f32(_ptr(out), _ptr(self.d_tgt[slot]), _ptr(self.d_itf[slot]), self.sb,
                                           self.sub.out_len, self.sub.L, _ptr(scores[lo:hi]), _stream()), "avz_sir_f32")
                if pcm:
                    _lib.check(lib.avz_f32_to_pcm16(_ptr(out), out.numel(), _ptr(self.w_out[slot]), _stream()), "avz_f32_to_pcm16")
                ev_cmp[slot] = self.s_cmp.record_event()
            with torch.cuda.stream(self.s_out):
                self.s_out.wait_event(ev_cmp[slot])
                h_out[lo:hi].copy_((self.w_out if pcm else self.d_out)[slot], non_blocking=True)
                ev_out[slot] = self.s_out.record_event()
        with torch.cuda.stream(self.s_cmp):
            if self.world > 1:
                import torch.distributed as dist
                dist.all_gather_into_tensor(self.scores_all[hs], scores)
            else:
                self.scores_all[hs].copy_(scores)
            ev_sc = self.s_cmp.record_event()
        with torch.cuda.stream(self.s_out):
            self.s_out.wait_event(ev_sc)
            self.h_scores[hs].copy_(self.scores_all[hs], non_blocking=True)
            self.done[hs] = self.s_out.record_event()
        return ticket

    def wait(self, ticket: int) -> Tuple[torch.Tensor, torch.Tensor]:
        """Block until step `ticket` is in pinned host memory -> (enhanced waveforms [B, n], all-gathered scores [B*world, 4])."""
        if not (self.n_submitted - self.host_slots <= ticket < self.n_submitted):
            raise _lib.AvzError("ticket is no longer (or not yet) held in a host slot")
        hs = ticket % self.host_slots
        self.done[hs].synchronize()
        return self.h_out[hs], self.h_scores[hs]

    def join(self) -> None:
        """Make the caller's stream wait for everything submitted so far."""
        cur = torch.cuda.current_stream()
        cur.wait_stream(self.s_cmp)
        cur.wait_stream(self.s_out)

    def run(self, mix_h: torch.Tensor, tgt_h: torch.Tensor, itf_h: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        return self.wait(self.submit(mix_h, tgt_h, itf_h))


def copy_ceiling(h2d_bytes: int, d2h_bytes: int, device, reps: int = 6) -> Dict[str, float]:
    """Bare-copy ceiling of the host<->device link for the end-to-end leg: per step one pinned host->device
    cudaMemcpyAsync of `h2d_bytes` and one device->host copy of `d2h_bytes`, nothing else running; `reps` steps back to
    back (sustained, like the leg itself - a single burst overstates what a shared host fabric holds when every rank
    copies at once: under torchrun all ranks run this at the same time).  Returns GB/s per direction, each direction
    alone and both at once, from the total time of the back-to-back steps."""
    h_in = torch.empty((h2d_bytes,), dtype=torch.uint8).pin_memory()
    h_out = torch.empty((d2h_bytes,), dtype=torch.uint8).pin_memory()
    d_in = torch.empty((h2d_bytes,), dtype=torch.uint8, device=device)
    d_out = torch.empty((d2h_bytes,), dtype=torch.uint8, device=device)
    s1, s2 = torch.cuda.Stream(device), torch.cuda.Stream(device)

    def run(h2d: bool, d2h: bool):
        torch.cuda.synchronize(device)
        a1, b1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a2, b2 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(s1):
            a1.record()
            for _ in range(reps if h2d else 0):
                d_in.copy_(h_in, non_blocking=True)
            b1.record()
        with torch.cuda.stream(s2):
            a2.record()
            for _ in range(reps if d2h else 0):
                h_out.copy_(d_out, non_blocking=True)
            b2.record()
        torch.cuda.synchronize(device)
        return a1.elapsed_time(b1) / reps, a2.elapsed_time(b2) / reps

    run(True, True)                                   # warm-up (first touch of the pinned pages)
    h_alone, _ = run(True, False)
    _, d_alone = run(False, True)
    h_both, d_both = run(True, True)
    return {"alone": {"h2d_GBps": h2d_bytes / (h_alone * 1e-3) / 1e9, "d2h_GBps": d2h_bytes / (d_alone * 1e-3) / 1e9,
                      "h2d_ms": h_alone, "d2h_ms": d_alone},
            "both_directions": {"h2d_GBps": h2d_bytes / (h_both * 1e-3) / 1e9, "d2h_GBps": d2h_bytes / (d_both * 1e-3) / 1e9,
                                "h2d_ms": h_both, "d2h_ms": d_both}}

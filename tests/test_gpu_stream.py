"""Streaming mask-MVDR: the GPU step kernel, fed hop by hop, against the float64 restatement of
the same recursion (oracle.streaming_mvdr).  The reference has no streaming mode: parity unpinned by construction."""
import numpy as np
import pytest
import torch

import oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def az():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    import avzoom
    avzoom._lib.load()
    return avzoom


def rel_l2(a, b):
    return float(np.linalg.norm(np.asarray(a, np.float64) - b) / (np.linalg.norm(b) + 1e-300))


@pytest.mark.parametrize("sigma,lam", [(1.0, 0.95), (1e-3, 0.9)])
def test_streaming_matches_recursion_oracle(az, sigma, lam):
    import dataclasses
    from avzoom import stream, synth
    S, dur = 3, 1.0
    mix, tgt, itf = synth.make_batch(5, S, dur, 2)
    L = mix.shape[-1]
    T = L // 128 + 1
    cfg = dataclasses.replace(az.PRESETS["baseline_oracle"], sigma=sigma, mic_dist=0.04)
    ocfg = O.PathConfig(sigma=sigma, mic_dist=0.04)
    rng = np.random.default_rng(0)
    # per-frame noise weights: the oracle IBM of the references, softened so that n_t never vanishes
    nw = np.stack([0.05 + 0.9 * O.ibm_noise_mask(O.stft_scipy(tgt[s], 512, 128), O.stft_scipy(itf[s], 512, 128))
                   for s in range(S)]).astype(np.float32)
    nw *= rng.uniform(0.5, 1.0, nw.shape).astype(np.float32)
    eng = stream.MvdrStream(S, cfg, lam=lam)
    out = eng.run(torch.from_numpy(mix).cuda(), torch.from_numpy(nw).cuda()).cpu().numpy()
    assert out.shape == (S, L)
    for s in range(S):
        it = iter(range(T))
        ref = O.streaming_mvdr(mix[s], lambda y, s=s, it=it: nw[s][:, next(it)], ocfg, lam=lam)
        assert ref.shape == (L,)
        assert rel_l2(out[s], ref) < 1e-4


def test_streaming_many_streams_and_latency_contract(az):
    """4096 concurrent streams (BASELINE config 4 shape): identical streams give identical outputs, the first three
    returned hops are the documented warm-up, and a silent stream stays exactly silent."""
    from avzoom import stream
    S = 4096
    eng = stream.MvdrStream(S, az.PRESETS["baseline_oracle"], lam=0.95)
    g = torch.Generator(device="cuda").manual_seed(0)
    base = torch.randn((1, 2, 128 * 12), device="cuda", generator=g) * 0.1
    x = base.repeat(S, 1, 1)
    x[7] = 0.0
    outs = []
    for h in range(12):
        outs.append(eng.step(x[:, :, h * 128:(h + 1) * 128].contiguous()).clone())
    y = torch.stack(outs, dim=1)             # [S, hops, 128]
    assert torch.equal(y[0], y[1]) and torch.equal(y[0], y[S - 1])
    assert float(y[7].abs().max()) == 0.0
    assert float(y[0, 0].abs().max()) == 0.0           # call 0 completes no frame; calls 1-2 return the trimmed-away
    assert float(y[0, 3:].abs().max()) > 0.0           # lead-in blocks (ignored by contract), call 3 the first output hop

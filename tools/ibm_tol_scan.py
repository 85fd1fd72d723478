"""Count IBM mismatches (float32 + fix-up vs all-float64) for the tolerance in AVZ_IBM_TOL (one process per value)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, avzoom
from avzoom import synth
cfg = avzoom.PRESETS["baseline_oracle"]
n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
mix, tgt, itf = synth.make_batch(2, n, 4.0, 3, start=7000, workers=8)
mix_d, tgt_d, itf_d = (torch.from_numpy(a).cuda() for a in (mix, tgt, itf))
bits, _, _ = avzoom.ibm_covariance(mix_d, tgt_d, itf_d, cfg)
ref = avzoom.ibm_exact_bits(tgt_d, itf_d, cfg)
x = (bits ^ ref)
nbad = int(sum(bin(int(v) & 0xffffffff).count("1") for v in x[x != 0].cpu().numpy()))
print("tol", os.environ.get("AVZ_IBM_TOL"), "bins", bits.shape[0] * bits.shape[1] * 257, "mismatching bins", nbad)
idx = torch.nonzero(x)
for b, t, w in idx.cpu().numpy()[:20]:
    d = int(x[b, t, w]) & 0xffffffff
    ks = [32 * w + i for i in range(32) if (d >> i) & 1]
    print("mismatch at b=%d t=%d bins=%s  fused=%d exact=%d" % (b, t, ks, (int(bits[b, t, w]) >> (ks[0] & 31)) & 1, (int(ref[b, t, w]) >> (ks[0] & 31)) & 1))
    # float64 spectra on the host for those bins
    import scipy.signal
    St = scipy.signal.stft(tgt[b].astype(np.float64), nperseg=512, noverlap=384)[2]
    Si = scipy.signal.stft(itf[b].astype(np.float64), nperseg=512, noverlap=384)[2]
    for k in ks:
        print("   k=%d |St|=%.6e |Si|=%.6e rel diff=%.3e frame rms=%.3e" % (k, abs(St[k, t]), abs(Si[k, t]), (abs(Si[k, t]) - abs(St[k, t])) / max(abs(St[k, t]), 1e-300), np.sqrt(np.mean(abs(St[:, t]) ** 2 + abs(Si[:, t]) ** 2))))

import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, avzoom, ctypes as C
from avzoom import synth, _lib
from avzoom.ops import _ptr, _stream, num_frames
B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
mix, tgt, itf = synth.make_batch(2, B, 4.0, 3, workers=8)
m, t, i = (torch.from_numpy(a).cuda() for a in (mix, tgt, itf))
lib = _lib.load()
L = mix.shape[-1]; T = num_frames(L, 512, 128)
nws = lib.avz_ibm_cov_ws_bytes(B, L, 512, 128)
ws = torch.zeros((nws,), dtype=torch.uint8, device="cuda")
bits = torch.empty((B, T, 9), dtype=torch.int32, device="cuda")
R = torch.empty((B, 257, 4), device="cuda"); ms = torch.empty((B, 257), device="cuda")
lib.avz_profile_enable(1)
buf = (C.c_float * 7)()
for _ in range(3):
    _lib.check(lib.avz_ibm_cov_f32(_ptr(m), _ptr(t), _ptr(i), B, L, 512, 128, 1e-6, _ptr(bits), _ptr(R), _ptr(ms), _ptr(ws), _stream()), "x")
    lib.avz_profile_get(buf, 7)
torch.cuda.synchronize()
w = ws.cpu().numpy()
cap = B * T * 8
found = None
for chunks in range(1, 200):
    pb = (B * chunks * 5 * 288 * 4 + 15) // 16 * 16
    if pb + 16 > len(w): break
    cnt = int(np.frombuffer(w[pb:pb + 4].tobytes(), dtype=np.uint32)[0])
    if 0 < cnt and pb + 16 + min(cnt, cap) * 8 <= len(w):
        ent = np.frombuffer(w[pb + 16:pb + 16 + min(cnt, 1000) * 8].tobytes(), dtype=np.uint64)
        if ((ent >> 32) < B).all() and (((ent >> 9) & 0x7fffff) < T).all() and ((ent & 511) <= 256).all():
            found = (chunks, cnt); break
print("B", B, "frames", B * T, "chunks,count", found, "entries/frame", found[1] / (B * T) if found else None, "cap", cap,
      "ibm ms %.3f fixup ms %.3f cov ms %.3f" % (buf[0], buf[1], buf[2]))

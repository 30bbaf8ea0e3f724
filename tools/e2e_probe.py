# development aid: wall clock of capy_sha3_batch_fixed on the headline shape (pinned host buffers)
import os as _os, sys as _sys
_sys.path.insert(0, _os.path.join(_os.path.dirname(_os.path.abspath(__file__)), '..'))
import hashlib, json, time
import numpy as np
from capycrypt_b200 import Engine
eng = Engine()
N = 1 << 20
h_in = eng.pinned(N * 64); h_out = eng.pinned(N * 32).reshape(N, 32)
h_in[:] = np.random.default_rng(1).integers(0, 256, N * 64, dtype=np.uint8)
for _ in range(5): eng.sha3_fixed(h_in, 64, 64, N, 256, out=h_out)
ts = []
for _ in range(5):
    t0 = time.perf_counter()
    for _ in range(50): eng.sha3_fixed(h_in, 64, 64, N, 256, out=h_out)
    ts.append((time.perf_counter() - t0) / 50)
ok = all(hashlib.sha3_256(h_in[64 * i:64 * i + 64].tobytes()).digest() == h_out[i].tobytes() for i in (0, 1, N // 2, N - 2, N - 1))
print(json.dumps({"ms_best": round(min(ts) * 1e3, 4), "ms_median": round(sorted(ts)[2] * 1e3, 4), "GBps_median": round(N * 64 / sorted(ts)[2] / 1e9, 2), "ok": ok}))

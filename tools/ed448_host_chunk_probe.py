# development aid: capy_ed448_fixed_base_batch / capy_ed448_sign_batch through pinned host buffers at sizes below and above
# the chunking threshold of the host entry points (csrc/ed448_api.cu: ed_chunk_count), wall clock per call
import os as _os, sys as _sys
_sys.path.insert(0, _os.path.join(_os.path.dirname(_os.path.abspath(__file__)), '..'))
import json, time
import numpy as np
from capycrypt_b200 import Engine
eng = Engine()
rs = np.random.default_rng(3)
def wall(fn, reps=3):
    fn()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    return (time.perf_counter() - t0) / reps * 1e3
for lg in (18, 19, 20):
    n = 1 << lg
    sc = eng.pinned(n * 56); sc[:] = rs.integers(0, 256, size=n * 56, dtype=np.uint8)
    out = eng.pinned(n * 112).reshape(n, 112)
    ms_fb = wall(lambda: eng.ed448_fixed_base(sc, out=out))
    pw = eng.pinned(n * 32); pw[:] = rs.integers(0, 256, size=n * 32, dtype=np.uint8)
    msg = eng.pinned(n * 256); msg[:] = rs.integers(0, 256, size=n * 256, dtype=np.uint8)
    po, mo = np.arange(n + 1, dtype=np.uint64) * 32, np.arange(n + 1, dtype=np.uint64) * 256
    h, z = eng.pinned(n * 56).reshape(n, 56), eng.pinned(n * 56).reshape(n, 56)
    ms_sign = wall(lambda: eng.ed448_sign(pw, po, msg, mo, 512, h_out=h, z_out=z))
    print(json.dumps({"log2_n": lg, "fixed_base_ms": round(ms_fb, 3), "fixed_base_Mps": round(n / ms_fb / 1e3, 2),
                      "sign_ms": round(ms_sign, 3), "sign_Mps": round(n / ms_sign / 1e3, 2)}))

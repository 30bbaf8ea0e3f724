# development aid: wall-clock of the host-buffer entry points on mid-size batches (looks for pathologies, not peaks)
import os as _os, sys as _sys
_sys.path.insert(0, _os.path.join(_os.path.dirname(_os.path.abspath(__file__)), '..'))
import json, time
import numpy as np
from capycrypt_b200 import Engine
eng = Engine(); rng = np.random.default_rng(3)
def t(fn, reps=3):
    fn(); best = 1e9
    for _ in range(reps):
        t0 = time.perf_counter(); fn(); best = min(best, time.perf_counter() - t0)
    return best
res = {}
n = 1 << 16
data = rng.integers(0, 256, size=n * 4096, dtype=np.uint8); off = np.arange(n + 1, dtype=np.uint64) * 4096
keys = rng.integers(0, 256, size=n * 32, dtype=np.uint8); koff = np.arange(n + 1, dtype=np.uint64) * 32
s = t(lambda: eng.kmac_xof(keys, koff, data, off, 512, b"My Tagged Application", 512)); res["kmac_2^16x4KB_GBps"] = round(n * 4096 / s / 1e9, 2)
s = t(lambda: eng.sha3(data, off, 512)); res["sha3_512_ragged_api_2^16x4KB_GBps"] = round(n * 4096 / s / 1e9, 2)
nonces = rng.integers(0, 256, size=n * 512, dtype=np.uint8)
s = t(lambda: eng.sponge_encrypt(keys, koff, nonces, 512, data, off, 512)); res["sponge_encrypt_2^16x4KB_GBps"] = round(n * 4096 / s / 1e9, 2)
m = 1 << 18
pw = rng.integers(0, 256, size=m * 32, dtype=np.uint8); po = np.arange(m + 1, dtype=np.uint64) * 32
msg = rng.integers(0, 256, size=m * 256, dtype=np.uint8); mo = np.arange(m + 1, dtype=np.uint64) * 256
s = t(lambda: eng.ed448_keygen(pw, po, 512), 2); res["keygen_2^18_Mps"] = round(m / s / 1e6, 2)
pub = eng.ed448_keygen(pw, po, 512)
s = t(lambda: eng.ed448_sign(pw, po, msg, mo, 512), 2); res["sign_2^18_Mps"] = round(m / s / 1e6, 2)
h, z = eng.ed448_sign(pw, po, msg, mo, 512)
s = t(lambda: eng.ed448_verify(pub, msg, mo, h, z, 512), 2); res["verify_2^18_Mps"] = round(m / s / 1e6, 2)
small = rng.integers(0, 256, size=(1 << 20) * 40, dtype=np.uint8); so = np.arange((1 << 20) + 1, dtype=np.uint64) * 40
s = t(lambda: eng.sha3(small, so, 256)); res["sha3_256_ragged_api_2^20x40B_Mmsgps"] = round((1 << 20) / s / 1e6, 2)
print(json.dumps(res))

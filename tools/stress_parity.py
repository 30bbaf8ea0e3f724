import sys, os
sys.path.insert(0, os.getcwd())
import numpy as np
from capycrypt_b200 import Engine, pack
from oracle import cpu
sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import test_fuzz_sponge_gpu as F
eng = Engine(); orc = cpu.get()
bad = 0
for seed in range(100, 160):
    rng = np.random.default_rng(seed)
    for d in (224, 256, 384, 512):
        n = int(rng.integers(1, 300))
        data, off = pack(F._items(rng, n))
        if not np.array_equal(eng.sha3(data, off, d), orc.sha3_batch(data, off, d, threads=0)): bad += 1; print("sha3 mismatch", seed, d)
        keys, koff = pack(F._items(rng, n, long_ok=False))
        custom = bytes(rng.integers(0, 256, size=int(rng.integers(0, 100)), dtype=np.uint8))
        ob = 8 * int(rng.integers(1, 600))
        if not np.array_equal(eng.kmac_xof(keys, koff, data, off, ob, custom, d), orc.kmac_xof_batch(keys, koff, data, off, ob, custom, d, threads=0)): bad += 1; print("kmac mismatch", seed, d)
# Ed448 stress
rng = np.random.default_rng(7)
for it in range(6):
    n = int(rng.integers(1, 600))
    pws, po = pack([bytes(rng.integers(0, 256, size=int(rng.integers(0, 50)), dtype=np.uint8)) for _ in range(n)])
    md, mo = pack([bytes(rng.integers(0, 256, size=int(rng.integers(0, 1500)), dtype=np.uint8)) for _ in range(n)])
    d = int(rng.choice([224, 256, 384, 512]))
    h, z = eng.ed448_sign(pws, po, md, mo, d)
    ho, zo = orc.sign_batch(pws, po, md, mo, d, threads=0)
    pub = eng.ed448_keygen(pws, po, d)
    if not (np.array_equal(h, ho) and np.array_equal(z, zo) and np.array_equal(pub, orc.keygen_batch(pws, po, d, threads=0))): bad += 1; print("ed448 mismatch", it)
    rc, ok = eng.ed448_verify(pub, md, mo, h, z, d)
    if rc or not ok.all(): bad += 1; print("verify fail", it)
    z2 = z.copy(); z2[::3, 55] ^= 1
    rc, ok2 = eng.ed448_verify(pub, md, mo, h, z2, d)
    if ok2[::3].any() or not ok2[1::3].all(): bad += 1; print("tamper detection", it)
print("stress done, problems:", bad)

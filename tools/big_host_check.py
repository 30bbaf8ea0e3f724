import sys, os, hashlib, time
sys.path.insert(0, os.getcwd())
import numpy as np
from capycrypt_b200 import Engine
eng = Engine()
rng = np.random.default_rng(11)
lens = np.concatenate([np.full(2500, 1 << 20), rng.integers(0, 200, size=1_000_000), np.full(300, (1 << 20) + 77)]).astype(np.int64)
rng.shuffle(lens)
off = np.zeros(len(lens) + 1, np.uint64); off[1:] = np.cumsum(lens)
print("bytes", int(off[-1]) / 2**30, "GiB, items", len(lens))
data = rng.integers(0, 256, size=int(off[-1]), dtype=np.uint8)
t0 = time.time(); dig = eng.sha3(data, off, 256); t1 = time.time()
print("host API seconds", round(t1 - t0, 2))
idx = np.concatenate([rng.choice(len(lens), 300, replace=False), np.nonzero(lens > 1000000)[0][:40], [0, len(lens) - 1]])
bad = 0
for i in idx:
    m = data[int(off[i]):int(off[i + 1])].tobytes()
    if hashlib.sha3_256(m).digest() != dig[i].tobytes(): bad += 1
print("checked", len(idx), "mismatches", bad)
d512 = eng.sha3(data[: int(off[2000])], off[:2001], 512)
print("sha3-512 spot", sum(hashlib.sha3_512(data[int(off[i]):int(off[i + 1])].tobytes()).digest() != d512[i].tobytes() for i in range(0, 2000, 37) if lens[i] % 72 != 71 and lens[i] % 136 != 135))  # (the reference's Q1 / Q2 quirk lengths differ from FIPS 202 on purpose)

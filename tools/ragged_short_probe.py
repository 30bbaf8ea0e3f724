# development aid: SHA3-256 over 2^20 short messages through the OFFSETS entry point (what the Rust shim calls) against the
# fixed-length entry point: equal 64-byte messages, and ragged 1..135-byte messages (every message is one boundary block)
import os as _os, sys as _sys
_sys.path.insert(0, _os.path.join(_os.path.dirname(_os.path.abspath(__file__)), '..'))
import json
import numpy as np, torch
from capycrypt_b200 import Engine
eng = Engine()
eng.set_plan_cache(True)
n = 1 << 20
g = torch.Generator(device="cuda"); g.manual_seed(1)
def best(fn, reps=5):
    fn(); torch.cuda.synchronize(); b = 1e9
    for _ in range(reps):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); b = min(b, e0.elapsed_time(e1))
    return b
data = torch.randint(0, 256, (n * 136,), dtype=torch.uint8, device="cuda", generator=g)
out = torch.zeros(n * 32, dtype=torch.uint8, device="cuda")
off64 = torch.arange(n + 1, dtype=torch.int64, device="cuda") * 64
res = {"fixed_64B_ms": round(best(lambda: eng.sha3_fixed_dev(data, 64, 64, n, 256, out)), 4),
       "offsets_64B_ms": round(best(lambda: eng.sha3_dev(data, off64, 256, out)), 4)}
lens = np.random.default_rng(2).integers(1, 136, size=n)
offr = torch.from_numpy(np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)).cuda()
res["offsets_ragged_1_135B_ms"] = round(best(lambda: eng.sha3_dev(data, offr, 256, out)), 4)
res["ragged_Mmsg_per_s"] = round(n / res["offsets_ragged_1_135B_ms"] / 1e3, 1)
print(json.dumps(res))

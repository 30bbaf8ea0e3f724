import os, sys, json
sys.path.insert(0, '.')
import numpy as np, torch
from capycrypt_b200 import Engine
eng = Engine()
def run(n):
    lens = np.full(n, 1 << 20, dtype=np.int64)
    off = np.zeros(n + 1, np.int64); off[1:] = np.cumsum(lens)
    data = torch.empty(int(off[-1]) + 16, dtype=torch.uint8, device="cuda"); data.random_(0, 256)
    t_off = torch.from_numpy(off).cuda(); out = torch.zeros(n * 64, dtype=torch.uint8, device="cuda")
    eng.sha3_dev(data, t_off, 512, out); torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); eng.sha3_dev(data, t_off, 512, out); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)
print(json.dumps({"lib": os.path.basename(os.environ.get("CAPY_GPU_LIB", "default")), "pair_4096x1MiB_ms": round(run(4096), 2), "warp_64x1MiB_ms": round(run(64), 2)}))

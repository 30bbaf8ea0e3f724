# quick device-resident throughput probe (development aid, not the contract bench)
import sys, time, json
import numpy as np, torch
from capycrypt_b200 import Engine
eng = Engine()
res = {}
def timeit(fn, reps=3):
    fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best
for logn in (16, 18, 20):
    n = 1 << logn
    sc = torch.randint(0, 256, (n * 56,), dtype=torch.uint8, device="cuda")
    out = torch.zeros(n * 112, dtype=torch.uint8, device="cuda")
    ms = timeit(lambda: eng.ed448_fixed_base_dev(sc, n, out))
    res[f"fixed_base_2^{logn}"] = {"ms": ms, "Mops": n / ms / 1e3}
    if logn <= 18:
        pts = out.clone()
        out2 = torch.zeros(n * 112, dtype=torch.uint8, device="cuda")
        ms = timeit(lambda: eng.ed448_var_base_dev(sc, pts, n, out2))
        res[f"var_base_2^{logn}"] = {"ms": ms, "Mops": n / ms / 1e3}
print(json.dumps(res, indent=1))

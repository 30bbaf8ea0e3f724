# development aid (not the contract bench): device-resident Ed448 throughput of one library variant.
# usage: CAPY_GPU_LIB=<path> python bench_ed448_probe.py
import os as _os, sys as _sys
_sys.path.insert(0, _os.path.join(_os.path.dirname(_os.path.abspath(__file__)), '..'))
import hashlib, json, os, sys
import numpy as np, torch
from capycrypt_b200 import Engine
eng = Engine()
def timeit(fn, reps=3):
    fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best
g = torch.Generator(device="cuda"); g.manual_seed(7)
res = {"lib": os.path.basename(os.environ.get("CAPY_GPU_LIB", "default"))}
n = 1 << 20
sc = torch.randint(0, 256, (n * 56,), dtype=torch.uint8, device="cuda", generator=g)
out = torch.zeros(n * 112, dtype=torch.uint8, device="cuda")
ms = timeit(lambda: eng.ed448_fixed_base_dev(sc, n, out))
res["fixed_2^20_ms"] = round(ms, 3); res["fixed_Mps"] = round(n / ms / 1e3, 2)
res["fixed_digest"] = hashlib.sha256(out[: 112 * 4096].cpu().numpy().tobytes()).hexdigest()[:12]
n2 = int(os.environ.get('VB_N', 1 << 18))
out2 = torch.zeros(n2 * 112, dtype=torch.uint8, device="cuda")
ms = timeit(lambda: eng.ed448_var_base_dev(sc, out, n2, out2), reps=2)
res["var_n"] = n2; res["var_2^18_ms"] = round(ms, 3); res["var_Mps"] = round(n2 / ms / 1e3, 2)
res["var_digest"] = hashlib.sha256(out2[: 112 * 4096].cpu().numpy().tobytes()).hexdigest()[:12]
print(json.dumps(res))
# verify pipeline (public inputs: direct table lookups), 2^17 items x 64 B messages
n3 = 1 << 17
pw = torch.randint(0, 256, (n3 * 16,), dtype=torch.uint8, device="cuda", generator=g)
pw_off = torch.arange(0, n3 + 1, dtype=torch.int64, device="cuda") * 16
msg = torch.randint(0, 256, (n3 * 64,), dtype=torch.uint8, device="cuda", generator=g)
msg_off = torch.arange(0, n3 + 1, dtype=torch.int64, device="cuda") * 64
pub = torch.zeros(n3 * 112, dtype=torch.uint8, device="cuda")
h = torch.zeros(n3 * 56, dtype=torch.uint8, device="cuda"); z = torch.zeros(n3 * 56, dtype=torch.uint8, device="cuda")
ok = torch.zeros(n3, dtype=torch.uint8, device="cuda")
eng.ed448_keygen_dev(pw, pw_off, 512, pub); eng.ed448_sign_dev(pw, pw_off, msg, msg_off, 512, h, z)
ms = timeit(lambda: eng.ed448_verify_dev(pub, msg, msg_off, h, z, 512, ok), reps=2)
res["verify_2^17_ms"] = round(ms, 3); res["verify_Mps"] = round(n3 / ms / 1e3, 2); res["verify_all_ok"] = bool(ok.all().item())
print(json.dumps(res))

import os, sys, json
sys.path.insert(0, '/root/repo' if os.path.isdir('/root/repo/capycrypt_b200') else '.')
import torch
from capycrypt_b200 import Engine
eng = Engine()
g = torch.Generator(device="cuda"); g.manual_seed(1)
n, mlen = 1 << 16, 4096
data = torch.randint(0, 256, (n * mlen,), dtype=torch.uint8, device="cuda", generator=g)
def best(fn, reps=5):
    fn(); torch.cuda.synchronize(); b = 1e9
    for _ in range(reps):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); b = min(b, e0.elapsed_time(e1))
    return b
for d in (256, 512):
    out = torch.zeros(n * d // 8, dtype=torch.uint8, device="cuda")
    fn = lambda: eng.sha3_fixed_dev(data, mlen, mlen, n, d, out)
    os.environ.pop("CAPY_NO_CHAIN_SPLIT", None); a = best(fn)
    os.environ["CAPY_NO_CHAIN_SPLIT"] = "1"; b = best(fn); os.environ.pop("CAPY_NO_CHAIN_SPLIT", None)
    rate = 136 if d == 256 else 72
    perms = (mlen + 1 + rate - 1) // rate
    print(json.dumps({"case": f"SHA3-{d} over 2^16 x 4 KB (fixed-length entry point)", "cut_ms": round(a, 4), "uncut_ms": round(b, 4),
                      "frac_lop3_cut": round(n * perms * 4320 / (a * 1e-3) / 18.47e12, 3), "frac_lop3_uncut": round(n * perms * 4320 / (b * 1e-3) / 18.47e12, 3)}))

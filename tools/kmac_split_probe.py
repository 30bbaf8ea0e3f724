# development aid: does cutting the chains of a uniform KMAC batch in half fill the partial wave?  Same bytes, same
# permutations: 2^16 x 4 KB against 2^17 x 2 KB and 2^18 x 1 KB (device-resident, CUDA events, best of 5)
import os as _os, sys as _sys
_sys.path.insert(0, _os.path.join(_os.path.dirname(_os.path.abspath(__file__)), '..'))
import json
import torch
from capycrypt_b200 import Engine
eng = Engine()
g = torch.Generator(device="cuda"); g.manual_seed(1)
total = (1 << 16) * 4096
data = torch.randint(0, 256, (total,), dtype=torch.uint8, device="cuda", generator=g)
keys = torch.randint(0, 256, ((1 << 18) * 32,), dtype=torch.uint8, device="cuda", generator=g)
def best(fn, reps=5):
    fn(); torch.cuda.synchronize(); b = 1e9
    for _ in range(reps):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); b = min(b, e0.elapsed_time(e1))
    return b
for lg in (16, 17, 18, 15):
    n = 1 << lg; mlen = total // n
    out = torch.zeros(n * 64, dtype=torch.uint8, device="cuda")
    ms = best(lambda: eng.kmac_xof_fixed_dev(keys, 32, 32, data, mlen, mlen, n, 512, b"My Tagged Application", 512, out))
    blocks = (136 + mlen + 3 + 135) // 136  # key block + message + trailer, prefix block cached
    print(json.dumps({"items": f"2^{lg}", "msg_bytes": mlen, "perms_per_item": blocks, "ms": round(ms, 4),
                      "warps_per_scheduler": round(n / 32 / 592, 2), "Gperm_per_s": round(n * blocks / ms / 1e6, 3)}))
# the engine's own cut (sponge_chain_kernel) against the uncut launch: BASELINE config 2 and the long-squeeze shapes
import os
n, mlen = 1 << 16, 4096
off = torch.arange(n + 1, dtype=torch.int64, device="cuda") * mlen
koff = torch.arange(n + 1, dtype=torch.int64, device="cuda") * 32
tag = torch.zeros(n * 64, dtype=torch.uint8, device="cuda")
big = torch.zeros(n * 4096, dtype=torch.uint8, device="cuda")
cases = {
    "kmacxof256 4 KB -> 64 B (cfg 2)": lambda: eng.kmac_xof_fixed_dev(keys, 32, 32, data, mlen, mlen, n, 512, b"My Tagged Application", 512, tag),
    "cshake256 4 KB -> 4 KB": lambda: eng.cshake_dev(data, off, 8 * 4096, b"", b"Email Signature", 512, big),
    "kmacxof256 4 KB -> 4 KB": lambda: eng.kmac_xof_dev(keys, koff, data, off, 8 * 4096, b"My Tagged Application", 512, big),
}
eng.set_plan_cache(True)
for name, fn in cases.items():
    os.environ.pop("CAPY_NO_CHAIN_SPLIT", None)
    a = best(fn)
    os.environ["CAPY_NO_CHAIN_SPLIT"] = "1"
    b = best(fn)
    os.environ.pop("CAPY_NO_CHAIN_SPLIT", None)
    print(json.dumps({"case": name, "cut_ms": round(a, 4), "uncut_ms": round(b, 4)}))

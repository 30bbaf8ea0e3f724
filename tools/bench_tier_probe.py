# development aid: chain-bound batches through the tiered kernel vs one thread per message (CAPY_FLAG_NO_PAIR)
import os as _os, sys as _sys
_sys.path.insert(0, _os.path.join(_os.path.dirname(_os.path.abspath(__file__)), '..'))
import json, os, sys
import numpy as np, torch
from capycrypt_b200 import Engine
eng = Engine()
dev = torch.device("cuda")
def run(lens, d, flags, data, reps=2):
    off = np.zeros(len(lens) + 1, np.int64); off[1:] = np.cumsum(lens)
    t_off = torch.from_numpy(off).to(dev)
    out = torch.zeros(len(lens) * d // 8, dtype=torch.uint8, device=dev)
    def call():
        eng._check(eng.lib.capy_sha3_batch_dev(eng._ctx, 0, eng._stream(), d, data.data_ptr(), t_off.data_ptr(), len(lens), out.data_ptr(), flags))
    call(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); call(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best, out.clone()
def mixed(total):
    rs = np.random.default_rng(5); lens, acc = [], 0
    while acc < total:
        c = np.exp(rs.uniform(np.log(64), np.log(1 << 20), size=8192)).astype(np.int64); lens.append(c); acc += int(c.sum())
    lens = np.concatenate(lens); return lens[: int(np.searchsorted(np.cumsum(lens), total)) + 1]
cases = [("mixed 16 GiB", mixed(16 << 30)), ("mixed 8 GiB", mixed(8 << 30)), ("mixed 4 GiB", mixed(4 << 30)), ("mixed 2 GiB", mixed(2 << 30)),
         ("4096 x 1 MiB (pair tier only: the warp-tier blocks would not all be resident)", np.full(4096, 1 << 20)),
         ("1024 x 1 MiB", np.full(1024, 1 << 20)), ("64 x 1 MiB", np.full(64, 1 << 20)), ("1 x 4 MiB", np.array([4 << 20]))]
for name, lens in cases:
    lens = np.asarray(lens, dtype=np.int64)
    data = torch.empty(int(lens.sum()) + 16, dtype=torch.uint8, device=dev); data.random_(0, 256)
    a, oa = run(lens, 512, 0, data)
    b, ob = run(lens, 512, 2, data)
    print(json.dumps({"case": name, "n": len(lens), "tiered_ms": round(a, 2), "one_thread_per_msg_ms": round(b, 2), "same": bool(torch.equal(oa, ob))}))
    del data
# warp-tier chain speed when c chains share a scheduler (flags = c << CAPY_FLAG_WARP_COSCHED_SHIFT forces the planner's
# co-scheduling factor): 74 SMs x 4 c messages of 1 MiB, and a full GPU (148 SMs) for c = 1..3 where the blocks fit
for c in (1, 2, 3):
    for sms in (74, 147):
        lens = np.full(sms * 4 * c, 1 << 20, dtype=np.int64)
        data = torch.empty(int(lens.sum()) + 16, dtype=torch.uint8, device=dev); data.random_(0, 256)
        a, oa = run(lens, 512, c << 8, data, reps=1)
        print(json.dumps({"case": f"warp tier, {c} chains per scheduler, {sms} SMs", "n": len(lens), "ms": round(a, 2),
                          "us_per_perm": round(a * 1e3 / 14564, 3)}))
        del data

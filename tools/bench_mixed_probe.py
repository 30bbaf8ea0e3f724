# development aid: BASELINE config 5 (mixed-size SHA3-512, log-uniform 64 B..1 MiB) device-resident probe.
# usage: python bench_mixed_probe.py [total_GiB]
import os as _os, sys as _sys
_sys.path.insert(0, _os.path.join(_os.path.dirname(_os.path.abspath(__file__)), '..'))
import json, sys
import numpy as np, torch
from capycrypt_b200 import Engine
total = float(sys.argv[1]) * (1 << 30) if len(sys.argv) > 1 else 4 * (1 << 30)
eng = Engine()
rnd = np.random.default_rng(5)
lens = []
s = 0
while s < total:
    chunk = np.exp(rnd.uniform(np.log(64), np.log(1 << 20), size=4096)).astype(np.int64)
    lens.append(chunk); s += int(chunk.sum())
lens = np.concatenate(lens)
cut = np.searchsorted(np.cumsum(lens), total) + 1
lens = lens[:cut]
off = np.zeros(len(lens) + 1, np.int64); off[1:] = np.cumsum(lens)
nbytes = int(off[-1])
data = torch.randint(0, 256, (nbytes + 8,), dtype=torch.uint8, device="cuda")
t_off = torch.from_numpy(off).cuda()
out = torch.zeros(len(lens) * 64, dtype=torch.uint8, device="cuda")
perms = int(((lens + 1 + 71) // 72).sum())
res = {"msgs": len(lens), "GiB": nbytes / 2**30, "perms": perms, "max_len": int(lens.max())}
for flags, name in ((0, "lpt_sorted"), (1, "no_sort")):
    def run():
        eng._check(eng.lib.capy_sha3_batch_dev(eng._ctx, 0, eng._stream(), 512, data.data_ptr(), t_off.data_ptr(), len(lens), out.data_ptr(), flags))
    run(); torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); run(); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    res[name] = {"ms": round(ms, 2), "GBps": round(nbytes / ms / 1e6, 1), "Gperm_s": round(perms / ms / 1e6, 3),
                 "frac_int_alu": round(perms * 4320 / (ms * 1e-3) / 18.5e12, 3)}
    if flags == 0: ref = out.clone()
    else: res["same_digests"] = bool(torch.equal(ref, out))
print(json.dumps(res))

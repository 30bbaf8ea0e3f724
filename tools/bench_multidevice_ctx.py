# Timing of the API the north star describes: ONE ctx over N devices (capy_gpu_init(devs, n)), host entry points that
# shard a batch over per-device worker threads and streams -- beside the process-per-GPU numbers of bench.py.
# usage: python tools/bench_multidevice_ctx.py [N ...]        (default: every N in 1, 2, 4, 8 that the box has)
import os as _os, sys as _sys
_sys.path.insert(0, _os.path.join(_os.path.dirname(_os.path.abspath(__file__)), '..'))
import hashlib, json, sys, time
import numpy as np, torch
from capycrypt_b200 import Engine

have = torch.cuda.device_count()
ns = [int(a) for a in sys.argv[1:]] or [n for n in (1, 2, 4, 8) if n <= have]
N_MSGS, MSG = 1 << 20, 64


def mixed(total, seed=5):
    rs = np.random.default_rng(seed); lens, acc = [], 0
    while acc < total:
        c = np.exp(rs.uniform(np.log(64), np.log(1 << 20), size=8192)).astype(np.int64); lens.append(c); acc += int(c.sum())
    lens = np.concatenate(lens); return lens[: int(np.searchsorted(np.cumsum(lens), total)) + 1]


for n in ns:
    if n > have:
        continue
    eng = Engine(devices=list(range(n)))
    # cfg 1, weak scaling like bench.py: n x 2^20 messages of 64 B in ONE call (pinned host buffers)
    tot = n * N_MSGS
    h_in, h_out = eng.pinned(tot * MSG), eng.pinned(tot * 32)
    rng = np.random.default_rng(1)
    blk = rng.integers(0, 256, size=N_MSGS * MSG, dtype=np.uint8)
    for k in range(n):
        h_in[k * len(blk):(k + 1) * len(blk)] = blk
    out2d = h_out.reshape(tot, 32)
    for _ in range(3):
        eng.sha3_fixed(h_in, MSG, MSG, tot, 256, out=out2d)
    reps = 20
    t0 = time.perf_counter()
    for _ in range(reps):
        eng.sha3_fixed(h_in, MSG, MSG, tot, 256, out=out2d)
    dt = (time.perf_counter() - t0) / reps
    assert h_out[:32].tobytes() == hashlib.sha3_256(blk[:MSG].tobytes()).digest()
    assert h_out[-32:].tobytes() == hashlib.sha3_256(blk[-MSG:].tobytes()).digest()
    link = max(eng.copy_probe(h_in[: N_MSGS * MSG], h_out[: N_MSGS * 32], reps=5, dev_index=k) for k in range(n))
    res = {"n_devices_in_ctx": n, "cfg1_e2e_GBps": tot * MSG / dt / 1e9, "cfg1_ms_per_call": dt * 1e3,
           "one_device_copy_probe_ms": link,
           "api": "capy_sha3_batch_fixed over one ctx (persistent per-device worker threads, 3 streams per device)"}
    # cfg 5, strong scaling: the 16 GiB mixed batch would need 16 GiB of pinned memory; a 4 GiB batch shows the same thing
    lens = mixed(4 << 30)
    off = np.zeros(len(lens) + 1, np.uint64); off[1:] = np.cumsum(lens)
    nb = int(off[-1])
    h5 = eng.pinned(nb)
    for p0 in range(0, nb, len(blk)):
        h5[p0:p0 + len(blk)] = blk[: min(len(blk), nb - p0)]
    eng.sha3(h5, off, 512)
    t0 = time.perf_counter()
    dg = eng.sha3(h5, off, 512)
    dt5 = time.perf_counter() - t0
    i = int(np.argmax(lens))
    assert dg[i].tobytes() != b"\0" * 64
    res.update({"cfg5_4GiB_e2e_GBps": nb / dt5 / 1e9, "cfg5_4GiB_ms": dt5 * 1e3, "cfg5_msgs": int(len(lens))})
    print(json.dumps(res), flush=True)
    eng.close()

"""Small run of every C entry point for compute-sanitizer (memcheck / racecheck):
    compute-sanitizer --tool memcheck python tests/host/sanitizer_driver.py
Sizes are tiny but cover the awkward shapes: empty items, unaligned starts, every load path of the sponge, the
three launch tiers, off-curve points, failed decryptions."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
from capycrypt_b200 import Engine, pack  # noqa: E402

eng = Engine()
rng = np.random.default_rng(7)


def items(lens):
    return [bytes(rng.integers(0, 256, size=int(n), dtype=np.uint8)) for n in lens]


lens = [0, 1, 7, 8, 9, 63, 64, 71, 72, 73, 135, 136, 137, 143, 144, 167, 168, 171, 172, 173, 300, 1000, 5000, 40000, 90001]
data, off = pack(items(lens))
for d in (224, 256, 384, 512):
    eng.sha3(data, off, d)
    eng.sha3_fixed(rng.integers(0, 256, size=100 * 80, dtype=np.uint8), 64, 80, 100, d)
    eng.sha3_fixed(rng.integers(0, 256, size=100 * 33, dtype=np.uint8), 33, 33, 100, d)
    keys, koff = pack(items([0, 5, 32, 200] * 6 + [64]))
    eng.kmac_xof(keys, koff, data, off, 512, b"custom", d)
    eng.kmac_xof(keys, koff, data, off, 8 * 333, b"", d)
    eng.cshake(data, off, 256, b"", b"", d)
    eng.cshake(data, off, 8 * 200, b"fn", b"s", d)
out_off = np.zeros(len(lens) + 1, np.uint64)
out_off[1:] = np.cumsum(lens[::-1])
eng.kmac_xof(keys, koff, data, off, 0, b"SKE", 512, out_off=out_off)
# long single message: warp tier; a few long + many short: all tiers
d1, o1 = pack(items([300000]))
eng.sha3(d1, o1, 512)
d2, o2 = pack(items([200000, 150000, 100000] + [50] * 2000))
eng.sha3(d2, o2, 256)
# Ed448
n = 40
pws, po = pack(items(rng.integers(0, 40, size=n)))
md, mo = pack(items(rng.integers(0, 700, size=n)))
pub = eng.ed448_keygen(pws, po, 512)
h, z = eng.ed448_sign(pws, po, md, mo, 512)
rc, ok = eng.ed448_verify(pub, md, mo, h, z, 512)
assert rc == 0 and ok.all()
bad = pub.copy()
bad[3, 0] ^= 1
eng.ed448_verify(bad, md, mo, h, z, 512)
sc = rng.integers(0, 256, size=n * 56, dtype=np.uint8)
eng.ed448_fixed_base(sc)
eng.ed448_var_base(sc, pub)
eng.ed448_var_base(sc, bad)
eng.ed448_ecdh(sc, pub)
# AE
nonces = rng.integers(0, 256, size=n * 512, dtype=np.uint8)
ct, tag = eng.sponge_encrypt(pws, po, nonces, 512, md, mo, 256)
out, ok = eng.sponge_decrypt(pws, po, nonces, 512, ct, mo, tag, 256)
assert ok.all() and np.array_equal(out, md)
tag2 = tag.copy()
tag2[::2, 0] ^= 1
eng.sponge_decrypt(pws, po, nonces, 512, ct, mo, tag2, 256)
rc, ct, tag, zp = eng.ed448_key_encrypt(pub, sc, md, mo, 512)
rc, out, ok = eng.ed448_key_decrypt(pws, po, zp, ct, mo, tag, 512)
assert rc == 0 and ok.all() and np.array_equal(out, md)
eng.ed448_key_decrypt(pws, po, bad, ct, mo, tag, 512)
eng.close()
print("sanitizer driver done")

"""N1 hand-off: when a maintainer has run rust/tests/dump_kats.rs next to the real crates and dropped
its output at tests/golden/ed448_crate_kat.json, these tests pin BOTH the oracle (CPU) and the engine (GPU, through the
C ABI) to capycrypt 0.7.5 + tiny_ed448_goldilocks 0.1.8 -- Ed448 parity goes from "unpinned" to "pinned".

Until that file exists the same checks run against tests/golden/ed448_golden.json (same schema: the frozen oracle
output), so the consumer below is exercised either way and the crate-file tests report themselves as skipped."""
import json
import os

import numpy as np
import pytest

from oracle import ref_ed448 as E
from oracle import ref_sha3 as R

H = bytes.fromhex
GOLD = os.path.join(os.path.dirname(__file__), "golden")
FILES = {"oracle-frozen": os.path.join(GOLD, "ed448_golden.json"), "crate": os.path.join(GOLD, "ed448_crate_kat.json")}


def _load(which):
    path = FILES[which]
    if not os.path.exists(path):
        pytest.skip("tests/golden/ed448_crate_kat.json not present: run rust/tests/dump_kats.rs with cargo (INTEGRATION.md)")
    with open(path) as f:
        return json.load(f)


def test_kat_inputs_match_the_frozen_fixtures():
    """rust/tests/dump_kats.rs reads tests/golden/ed448_kat_inputs.json: it must describe the same cases, in the same
    order, as ed448_golden.json, so that a crate dump is comparable record by record."""
    g = _load("oracle-frozen")
    inp = json.load(open(os.path.join(GOLD, "ed448_kat_inputs.json")))
    assert len(inp["cases"]) == len(g["keygen"]) == len(g["sign"]) == len(g["key_encrypt"]) == len(g["sha3_encrypt"])
    for c, kg, sg, ke, se in zip(inp["cases"], g["keygen"], g["sign"], g["key_encrypt"], g["sha3_encrypt"]):
        assert (c["d"], c["pw"]) == (kg["d"], kg["pw"]) == (sg["d"], sg["pw"])
        assert c["msg"] == sg["msg"] == ke["msg"] == se["msg"] and c["k_rand"] == ke["k_rand"] and c["nonce"] == se["nonce"]


@pytest.mark.parametrize("which", ["oracle-frozen", "crate"])
def test_oracle_matches(which):
    g = _load(which)
    for v in g["keygen"]:
        assert E.point_to_bytes(E.keygen(H(v["pw"]), v["d"])).hex() == v["pub_xy"], ("keygen", v["d"], v["pw"])
    for v in g["sign"]:
        h, z = E.sign(H(v["pw"]), H(v["msg"]), v["d"])
        assert (h.hex(), z.hex()) == (v["h"], v["z"]), ("sign", v["d"], v["pw"])
    for v in g["key_encrypt"][::2]:
        pub = E.keygen(H(v["pw"]), v["d"])
        ct, tag, zp = E.key_encrypt(pub, H(v["msg"]), v["d"], H(v["k_rand"]))
        assert (ct.hex(), tag.hex(), E.point_to_bytes(zp).hex()) == (v["ct"], v["tag"], v["z_xy"]), ("key_encrypt", v["d"])
    for v in g["sha3_encrypt"][::2]:
        ct, tag = R.sha3_encrypt(H(v["msg"]), H(v["pw"]), v["d"], H(v["nonce"]))
        assert (ct.hex(), tag.hex()) == (v["ct"], v["tag"])
    for v in g.get("rfc8032_xy", []):  # crate dump only: [s]G for the RFC 8032 scalars = the generator question (App. C.4)
        k = int.from_bytes(H(v["scalar_be56"]), "big")
        assert E.point_to_bytes(E.scalar_mult(k % E.R, E.GENERATOR)).hex() == v["xy"], "the crate's generator is not the RFC 8032 point"


@pytest.mark.gpu
@pytest.mark.parametrize("which", ["oracle-frozen", "crate"])
def test_engine_matches(which, engine):
    from capycrypt_b200 import pack

    g = _load(which)
    for d in (224, 256, 384, 512):
        kg = [v for v in g["keygen"] if v["d"] == d]
        pws, po = pack([H(v["pw"]) for v in kg])
        pub = engine.ed448_keygen(pws, po, d)
        assert [r.tobytes().hex() for r in pub] == [v["pub_xy"] for v in kg]
        sg = [v for v in g["sign"] if v["d"] == d]
        pws, po = pack([H(v["pw"]) for v in sg])
        md, mo = pack([H(v["msg"]) for v in sg])
        h, z = engine.ed448_sign(pws, po, md, mo, d)
        assert [r.tobytes().hex() for r in h] == [v["h"] for v in sg] and [r.tobytes().hex() for r in z] == [v["z"] for v in sg]
        spub = engine.ed448_keygen(pws, po, d)
        rc, ok = engine.ed448_verify(spub, md, mo, h, z, d)
        assert rc == 0 and ok.all()
        ke = [v for v in g["key_encrypt"] if v["d"] == d]
        pws, po = pack([H(v["pw"]) for v in ke])
        kpub = engine.ed448_keygen(pws, po, d)
        md, mo = pack([H(v["msg"]) for v in ke])
        rc, ct, tag, zz = engine.ed448_key_encrypt(kpub, b"".join(H(v["k_rand"]) for v in ke), md, mo, d)
        assert rc == 0
        for i, v in enumerate(ke):
            assert ct[int(mo[i]):int(mo[i + 1])].tobytes().hex() == v["ct"]
            assert tag[i].tobytes().hex() == v["tag"] and zz[i].tobytes().hex() == v["z_xy"]
        se = [v for v in g["sha3_encrypt"] if v["d"] == d]
        pws, po = pack([H(v["pw"]) for v in se])
        md, mo = pack([H(v["msg"]) for v in se])
        ct, tag = engine.sponge_encrypt(pws, po, b"".join(H(v["nonce"]) for v in se), 512, md, mo, d)
        for i, v in enumerate(se):
            assert ct[int(mo[i]):int(mo[i + 1])].tobytes().hex() == v["ct"] and tag[i].tobytes().hex() == v["tag"]
    if "rfc8032_xy" in g:
        sc = np.frombuffer(b"".join(H(v["scalar_be56"]) for v in g["rfc8032_xy"]), dtype=np.uint8)
        got = engine.ed448_fixed_base(sc)
        assert [r.tobytes().hex() for r in got] == [v["xy"] for v in g["rfc8032_xy"]]

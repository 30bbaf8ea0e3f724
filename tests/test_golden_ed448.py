"""Committed golden fixtures for the Ed448 / AE side (tests/golden/ed448_golden.json, made by
tests/golden/make_ed448_golden.py): RFC 8032 7.4 key pairs (absolute pin of the base point and of [s]G) and frozen
outputs of the restated reference (KeyPair::new ecc/keypair.rs:41-51, sign ecc/signable.rs:40-57, key_encrypt
ecc/encryptable.rs:34-50, sha3_encrypt sha3/encryptable.rs:29-45).

CPU part: the Python oracle and the C oracle reproduce the fixtures.  GPU part: the engine does, through the C ABI."""
import json
import os

import numpy as np
import pytest

from oracle import ref_ed448 as E
from oracle import ref_sha3 as R

H = bytes.fromhex


@pytest.fixture(scope="module")
def gold():
    with open(os.path.join(os.path.dirname(__file__), "golden", "ed448_golden.json")) as f:
        return json.load(f)


# ---- CPU ---------------------------------------------------------------------------------------------------
def test_oracle_reproduces_rfc8032(gold, oracle):
    for v in gold["rfc8032"]:
        k = int.from_bytes(H(v["scalar_be56"]), "big")
        assert E.rfc8032_encode(E.scalar_mult(k % E.R, E.GENERATOR)).hex() == v["public"]
        xy = oracle.fixed_base_batch(np.frombuffer(H(v["scalar_be56"]), dtype=np.uint8), threads=1)[0].tobytes()
        assert E.rfc8032_encode(E.point_from_bytes(xy)).hex() == v["public"]


def test_oracle_reproduces_fixtures(gold):
    for v in gold["keygen"][::3]:
        assert E.point_to_bytes(E.keygen(H(v["pw"]), v["d"])).hex() == v["pub_xy"]
    for v in gold["sign"][::3]:
        h, z = E.sign(H(v["pw"]), H(v["msg"]), v["d"])
        assert (h.hex(), z.hex()) == (v["h"], v["z"])
    for v in gold["sha3_encrypt"]:
        ct, tag = R.sha3_encrypt(H(v["msg"]), H(v["pw"]), v["d"], H(v["nonce"]))
        assert (ct.hex(), tag.hex()) == (v["ct"], v["tag"])


# ---- GPU ---------------------------------------------------------------------------------------------------
def _pack(items):
    from capycrypt_b200 import pack

    return pack(items)


@pytest.mark.gpu
def test_engine_reproduces_rfc8032(gold, engine):
    sc = np.frombuffer(b"".join(H(v["scalar_be56"]) for v in gold["rfc8032"]), dtype=np.uint8)
    got = engine.ed448_fixed_base(sc)
    for v, row in zip(gold["rfc8032"], got):
        assert E.rfc8032_encode(E.point_from_bytes(row.tobytes())).hex() == v["public"]


@pytest.mark.gpu
def test_engine_reproduces_keygen_sign_fixtures(gold, engine):
    for d in (224, 256, 384, 512):
        kg = [v for v in gold["keygen"] if v["d"] == d]
        pws, po = _pack([H(v["pw"]) for v in kg])
        pub = engine.ed448_keygen(pws, po, d)
        assert [r.tobytes().hex() for r in pub] == [v["pub_xy"] for v in kg]
        sg = [v for v in gold["sign"] if v["d"] == d]
        pws, po = _pack([H(v["pw"]) for v in sg])
        md, mo = _pack([H(v["msg"]) for v in sg])
        h, z = engine.ed448_sign(pws, po, md, mo, d)
        assert [r.tobytes().hex() for r in h] == [v["h"] for v in sg]
        assert [r.tobytes().hex() for r in z] == [v["z"] for v in sg]
        rc, ok = engine.ed448_verify(pub, md, mo, h, z, d)
        assert rc == 0 and ok.all()


@pytest.mark.gpu
def test_engine_reproduces_encryption_fixtures(gold, engine):
    for d in (224, 256, 384, 512):
        ke = [v for v in gold["key_encrypt"] if v["d"] == d]
        pub = b"".join(E.point_to_bytes(E.keygen(H(v["pw"]), d)) for v in ke)
        md, mo = _pack([H(v["msg"]) for v in ke])
        rc, ct, tag, z = engine.ed448_key_encrypt(pub, b"".join(H(v["k_rand"]) for v in ke), md, mo, d)
        assert rc == 0
        for i, v in enumerate(ke):
            assert ct[int(mo[i]):int(mo[i + 1])].tobytes().hex() == v["ct"]
            assert tag[i].tobytes().hex() == v["tag"] and z[i].tobytes().hex() == v["z_xy"]
        se = [v for v in gold["sha3_encrypt"] if v["d"] == d]
        pws, po = _pack([H(v["pw"]) for v in se])
        md, mo = _pack([H(v["msg"]) for v in se])
        ct, tag = engine.sponge_encrypt(pws, po, b"".join(H(v["nonce"]) for v in se), 512, md, mo, d)
        for i, v in enumerate(se):
            assert ct[int(mo[i]):int(mo[i + 1])].tobytes().hex() == v["ct"] and tag[i].tobytes().hex() == v["tag"]

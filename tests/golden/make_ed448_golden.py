"""Generates tests/golden/ed448_golden.json.

The reference holds no absolute Ed448 / AE vector (its tests are round trips) and its arithmetic crate
is not in the reference tree, so the fixtures come from two sources:
  * rfc8032: the Ed448 (secret key, public key) test vectors of RFC 8032 section 7.4, re-checked against OpenSSL
    when this script runs -- an absolute pin of the base point and of [s]G;
  * oracle/ref_ed448.py + oracle/ref_sha3.py (the literal restatements of src/ecc/*.rs and src/sha3/*.rs) run on fixed
    inputs: KeyPair::new, sign, key_encrypt, sha3_encrypt.  They freeze today's oracle output so that neither the
    oracle nor the engine can drift unnoticed.
usage: python tests/golden/make_ed448_golden.py
"""
import hashlib
import json
import os
import random
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_ed448 as E  # noqa: E402
from oracle import ref_sha3 as R  # noqa: E402

RFC8032 = [  # RFC 8032 7.4: "-----Blank", "-----1 octet", "-----1 octet (with context)" share keys 2/3
    ("6c82a562cb808d10d632be89c8513ebf6c929f34ddfa8c9f63c9960ef6e348a3528c8a3fcc2f044e39a3fc5b94492f8f032e7549a20098f95b",
     "5fd7449b59b461fd2ce787ec616ad46a1da1342485a70e1f8a0ea75d80e96778edf124769b46c7061bd6783df1e50f6cd1fa1abeafe8256180"),
    ("c4eab05d357007c632f3dbb48489924d552b08fe0c353a0d4a1f00acda2c463afbea67c5e8d2877c5e3bc397a659949ef8021e954e0a12274e",
     "43ba28f430cdff456ae531545f7ecd0ac834a55d9358c0372bfa0c6c6798c0866aea01eb00742802b8438ea4cb82169c235160627b4c3a9480"),
]


def rfc_scalar(sk: bytes) -> int:
    h = bytearray(hashlib.shake_256(sk).digest(114)[:57])
    h[0] &= 0xFC
    h[55] |= 0x80
    h[56] = 0
    return int.from_bytes(h, "little")


def main():
    out = {"rfc8032": [], "keygen": [], "sign": [], "key_encrypt": [], "sha3_encrypt": []}
    try:
        from cryptography.hazmat.primitives import serialization as S
        from cryptography.hazmat.primitives.asymmetric.ed448 import Ed448PrivateKey
    except ImportError:
        Ed448PrivateKey = None
    for sk_hex, pk_hex in RFC8032:
        sk = bytes.fromhex(sk_hex)
        if Ed448PrivateKey:
            pk = Ed448PrivateKey.from_private_bytes(sk).public_key().public_bytes(S.Encoding.Raw, S.PublicFormat.Raw)
            assert pk.hex() == pk_hex, "RFC 8032 vector does not match OpenSSL"
        s = rfc_scalar(sk)
        assert E.rfc8032_encode(E.scalar_mult(s % E.R, E.GENERATOR)).hex() == pk_hex, "oracle disagrees with RFC 8032"
        out["rfc8032"].append({"secret": sk_hex, "scalar_be56": s.to_bytes(56, "big").hex(), "public": pk_hex})
    rnd = random.Random(20261018)
    for d in (224, 256, 384, 512):
        for pw_len, msg_len in ((0, 0), (5, 17), (32, 136), (64, 300)):
            pw, msg = rnd.randbytes(pw_len), rnd.randbytes(msg_len)
            pub = E.keygen(pw, d)
            out["keygen"].append({"d": d, "pw": pw.hex(), "pub_xy": E.point_to_bytes(pub).hex()})
            h, z = E.sign(pw, msg, d)
            assert E.verify(pub, msg, h, z, d)
            out["sign"].append({"d": d, "pw": pw.hex(), "msg": msg.hex(), "h": h.hex(), "z": z.hex()})
            k = rnd.randbytes(56)
            ct, tag, zp = E.key_encrypt(pub, msg, d, k)
            ok, pt = E.key_decrypt(pw, ct, d, zp, tag)
            assert ok and pt == msg
            out["key_encrypt"].append({"d": d, "pw": pw.hex(), "k_rand": k.hex(), "msg": msg.hex(), "ct": ct.hex(),
                                       "tag": tag.hex(), "z_xy": E.point_to_bytes(zp).hex()})
            nonce = rnd.randbytes(512)
            ct, tag = R.sha3_encrypt(msg, pw, d, nonce)
            out["sha3_encrypt"].append({"d": d, "pw": pw.hex(), "nonce": nonce.hex(), "msg": msg.hex(), "ct": ct.hex(),
                                        "tag": tag.hex()})
    path = os.path.join(ROOT, "tests", "golden", "ed448_golden.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=1)
    print(path, {k: len(v) for k, v in out.items()})


if __name__ == "__main__":
    main()

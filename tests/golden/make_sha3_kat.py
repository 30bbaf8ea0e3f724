#!/usr/bin/env python3
"""Harvest the known-answer vectors of capyCRYPT's own SHA3-path tests into sha3_kat.json.

Run in the build container only (it reads /root/reference, which does not exist on the
GPU box):   python tests/golden/make_sha3_kat.py
Sources of the vectors:
  src/sha3/shake_functions.rs:92-288   SHA3-224/256/384/512, tagged hash, cSHAKE, KMACXOF
  src/sha3/sponge.rs:99-190            byte_pad / left_encode / right_encode
  tests/integration_tests.rs:84-93     SHA3-256("")
Only test DATA (byte arrays the reference asserts on) is extracted; no code is copied.
"""
import json
import os
import re

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def rust_fn_bodies(path):
    src = open(path).read()
    out = {}
    for m in re.finditer(r"fn (test_\w+)\(\) \{", src):
        start = m.end()
        depth, i = 1, start
        while depth:
            depth += {"{": 1, "}": -1}.get(src[i], 0)
            i += 1
        out[m.group(1)] = src[start:i - 1]
    return out


def byte_arrays(body):
    """All `[0x.., ..]` / `[1, 2, ..]` array literals in a test body, in order."""
    arrs = []
    for m in re.finditer(r"=\s*\[([0-9a-fA-Fx,\s]+)\];", body):
        toks = [t for t in re.split(r"[,\s]+", m.group(1)) if t]
        arrs.append(bytes(int(t, 0) for t in toks))
    return arrs


def main():
    sf = rust_fn_bodies(f"{REF}/src/sha3/shake_functions.rs")
    sp = rust_fn_bodies(f"{REF}/src/sha3/sponge.rs")
    it = rust_fn_bodies(f"{REF}/tests/integration_tests.rs")
    nist200 = bytes(range(200))  # NIST_DATA_SPONGE_INIT, src/sha3/constants.rs:6-20
    kat = {"sha3": [], "tagged_hash": [], "cshake": [], "kmac_xof": [], "encoders": {}}

    for d, fn in ((224, "test_shake_224"), (256, "test_shake_256"), (384, "test_shake_384")):
        a = byte_arrays(sf[fn])
        kat["sha3"].append({"d": d, "msg": "", "digest": a[0].hex(), "src": f"shake_functions.rs:{fn}"})
        kat["sha3"].append({"d": d, "msg": b"test".hex(), "digest": a[1].hex(), "src": f"shake_functions.rs:{fn}"})
    a = byte_arrays(sf["test_shake_512"])
    kat["sha3"].append({"d": 512, "msg": b"test".hex(), "digest": a[0].hex(), "src": "shake_functions.rs:test_shake_512"})
    hx = re.search(r'expected = "([0-9a-f]{64})"', it["test_hashable"]).group(1)
    kat["sha3"].append({"d": 256, "msg": "", "digest": hx, "src": "tests/integration_tests.rs:test_hashable"})

    a = byte_arrays(sf["test_compute_tagged_hash_256"])
    kat["tagged_hash"].append({"d": 256, "pw": "", "s": "", "msg": "", "digest": a[0].hex()})
    a = byte_arrays(sf["test_compute_tagged_hash_512"])
    kat["tagged_hash"].append({"d": 512, "pw": b"test".hex(), "s": "", "msg": "", "digest": a[0].hex()})

    a = byte_arrays(sf["test_cshake_256"])
    kat["cshake"].append({"d": 256, "x": nist200.hex(), "l": 256, "n": "", "s": b"Email Signature".hex(), "out": a[0].hex()})
    a = byte_arrays(sf["test_cshake_512"])
    kat["cshake"].append({"d": 512, "x": nist200.hex(), "l": 512, "n": "", "s": b"Email Signature".hex(), "out": a[0].hex()})

    a = byte_arrays(sf["test_kmac_256"])
    hx = re.search(r'expected = "([0-9a-f]+)"', sf["test_kmac_256"]).group(1)
    kat["kmac_xof"].append({"d": 512, "k": a[0].hex(), "x": "00010203", "l": 64,
                            "s": b"My Tagged Application".hex(), "out": hx})
    a = byte_arrays(sf["test_kmac_512"])
    kat["kmac_xof"].append({"d": 512, "k": a[0].hex(), "x": nist200.hex(), "l": 512,
                            "s": b"My Tagged Application".hex(), "out": a[1].hex()})

    # encoders: (value -> expected bytes) pairs asserted in sponge.rs tests
    def enc_pairs(body):
        vals = [int(v.replace("_", ""), 0) for v in re.findall(r"let val = (0x[0-9A-Fa-f]+|\d+);", body)]
        exps = [bytes(int(t, 0) for t in re.split(r"[,\s]+", m) if t)
                for m in re.findall(r"let expected = \[([0-9,\s]+)\];", body)]
        pairs = [{"value": str(v), "out": e.hex()} for v, e in zip(vals, exps)]
        pairs.append({"value": "200", "out": exps[len(vals)].hex()})  # val_len of NIST data
        return pairs

    kat["encoders"]["left_encode"] = enc_pairs(sp["test_left_encode"])
    kat["encoders"]["right_encode"] = enc_pairs(sp["test_right_encode"])
    exps = [bytes(int(t, 0) for t in re.split(r"[,\s]+", m) if t)
            for m in re.findall(r"let expected = \[([0-9,\s]+)\];", sp["test_bytepad"])]
    kat["encoders"]["byte_pad"] = [
        {"input": b"test".hex(), "w": 4, "out": exps[0].hex()},
        {"input": nist200.hex(), "w": 200, "out": exps[1].hex()},
    ]

    with open(os.path.join(HERE, "sha3_kat.json"), "w") as f:
        json.dump(kat, f, indent=1)
    n = sum(len(v) for k, v in kat.items() if k != "encoders") + sum(len(v) for v in kat["encoders"].values())
    print(f"wrote sha3_kat.json with {n} vectors")


if __name__ == "__main__":
    main()

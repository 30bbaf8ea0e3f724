"""CPU suite: the oracle (Python and C restatements) against every known-answer vector the
reference's own tests hold for the SHA3 path (tests/golden/sha3_kat.json, harvested from
src/sha3/shake_functions.rs:92-288, src/sha3/sponge.rs:99-190, tests/integration_tests.rs:84-93)."""
import hashlib
import random

import pytest

from oracle import ref_sha3 as R

H = bytes.fromhex


def test_py_sha3_kat(kat):
    for v in kat["sha3"]:
        assert R.compute_sha3_hash(H(v["msg"]), v["d"]).hex() == v["digest"], v["src"]


def test_py_tagged_hash_kat(kat):
    for v in kat["tagged_hash"]:
        assert R.compute_tagged_hash(H(v["msg"]), H(v["pw"]), H(v["s"]), v["d"]).hex() == v["digest"]


def test_py_cshake_kmac_kat(kat):
    for v in kat["cshake"]:
        assert R.cshake(H(v["x"]), v["l"], H(v["n"]), H(v["s"]), v["d"]).hex() == v["out"]
    for v in kat["kmac_xof"]:
        assert R.kmac_xof(H(v["k"]), H(v["x"]), v["l"], H(v["s"]), v["d"]).hex() == v["out"]


def test_py_encoders_kat(kat):
    for v in kat["encoders"]["left_encode"]:
        assert R.left_encode(int(v["value"])).hex() == v["out"]
    for v in kat["encoders"]["right_encode"]:
        assert R.right_encode(int(v["value"])).hex() == v["out"]
    for v in kat["encoders"]["byte_pad"]:
        assert R.byte_pad(H(v["input"]), v["w"]).hex() == v["out"]


def test_c_oracle_kat(kat, oracle):
    for v in kat["sha3"]:
        assert oracle.sha3(H(v["msg"]), v["d"]).hex() == v["digest"]
    for v in kat["tagged_hash"]:
        assert oracle.kmac_xof(H(v["pw"]), H(v["msg"]), v["d"], H(v["s"]), v["d"]).hex() == v["digest"]
    for v in kat["cshake"]:
        assert oracle.cshake(H(v["x"]), v["l"], H(v["n"]), H(v["s"]), v["d"]).hex() == v["out"]
    for v in kat["kmac_xof"]:
        assert oracle.kmac_xof(H(v["k"]), H(v["x"]), v["l"], H(v["s"]), v["d"]).hex() == v["out"]


def test_invalid_secparam():
    with pytest.raises(ValueError):
        R.compute_sha3_hash(b"", 128)
    with pytest.raises(ValueError):
        R.kmac_xof(b"", b"", 256, b"", 300)


def test_sha3_256_is_fips_for_all_lengths():
    """Quirk Q2: D256 is the one parameter where the reference is FIPS-exact everywhere."""
    rnd = random.Random(7)
    for n in list(range(0, 420)) + [1000, 4096, 4097]:
        m = rnd.randbytes(n)
        assert R.compute_sha3_hash(m, 256) == hashlib.sha3_256(m).digest()


def test_fips_divergence_sets():
    """the exact lengths (< 600) where the reference departs from FIPS 202."""
    expect = {
        512: [71, 135, 143, 215, 271, 287, 359, 407, 431, 503, 543, 575],
        384: [103, 135, 207, 271, 311, 407, 415, 519, 543],
        224: [135, 143, 271, 287, 407, 431, 543, 575],
        256: [],
    }
    fips = {224: hashlib.sha3_224, 256: hashlib.sha3_256, 384: hashlib.sha3_384, 512: hashlib.sha3_512}
    for d, want in expect.items():
        got = [n for n in range(600) if R.compute_sha3_hash(bytes(n), d) != fips[d](bytes(n)).digest()]
        assert got == want, d


def test_c_matches_py_sha3_sweep(oracle):
    rnd = random.Random(1)
    for d in (224, 256, 384, 512):
        for n in list(range(0, 300)) + [rnd.randrange(300, 3000) for _ in range(10)]:
            m = rnd.randbytes(n)
            assert oracle.sha3(m, d) == R.compute_sha3_hash(m, d), (d, n)


def test_c_matches_py_kmac_cshake_sweep(oracle):
    rnd = random.Random(2)
    for d in (224, 256, 384, 512):
        for n in list(range(0, 200, 5)) + [rnd.randrange(300, 1500) for _ in range(6)]:
            m = rnd.randbytes(n)
            key = rnd.randbytes(rnd.choice([0, 1, 32, 56, 131, 132, 167, rnd.randrange(0, 300)]))
            s = rnd.randbytes(rnd.randrange(0, 40))
            l = rnd.choice([8, 64, 448, 512, 1024, 2000, 8 * 400])
            assert oracle.kmac_xof(key, m, l, s, d) == R.kmac_xof(key, m, l, s, d), (d, n)
            nn = rnd.choice([b"", b"KMAC", b"abc"])
            s2 = rnd.choice([b"", s])
            assert oracle.cshake(m, l, nn, s2, d) == R.cshake(m, l, nn, s2, d), (d, n, nn, s2)


def test_fips_shake_extra_vs_hashlib():
    rnd = random.Random(3)
    for n in list(range(0, 300, 7)) + [135, 136, 137, 167, 168, 169]:
        m = rnd.randbytes(n)
        assert R.fips_shake(m, 100, 256) == hashlib.shake_256(m).digest(100)
        assert R.fips_shake(m, 200, 128) == hashlib.shake_128(m).digest(200)


def test_sponge_ae_roundtrip_and_restore():
    """tests/integration_tests.rs:250-262 restated: failed decrypt leaves the ciphertext intact."""
    rnd = random.Random(4)
    msg, pw, z = rnd.randbytes(1000), rnd.randbytes(16), rnd.randbytes(512)
    ct, tag = R.sha3_encrypt(msg, pw, 512, z)
    ok, pt = R.sha3_decrypt(ct, pw, 512, z, tag)
    assert ok and pt == msg
    ok, buf = R.sha3_decrypt(ct, b"wrong", 512, z, tag)
    assert not ok and buf == ct

"""CPU suite: pins the Ed448 oracle (which is 'parity unpinned' against the un-vendored Rust crate
tiny_ed448_goldilocks 0.1.8) with independent evidence:
  * OpenSSL (cryptography): RFC 8032 Ed448 public keys and RFC 7748 X448 shared secrets;
  * group laws; the reference's own self-consistency tests (tests/integration_tests.rs:20-130) restated;
  * C restatement == Python restatement on seeded inputs."""
import hashlib
import random

import numpy as np
import pytest

from oracle import cpu
from oracle import ref_ed448 as E


def test_curve_constants():
    assert E.on_curve(E.GENERATOR)
    assert E.scalar_mult(E.R, E.GENERATOR) == E.IDENTITY
    assert E.R % 4 == 3 and E.P % 4 == 3
    # the alternative generator of SURVEY.md App. C.4 (y = -3) is also on the curve with order r
    alt_x = 0x29C4D0C4EC185FD7B71AEB57B0627B79758FB15699CA3841492BDB973652ECB3977DCD53742C2095EF3967A7091607D98B7BD2308823FC56
    alt = (alt_x, (-3) % E.P)
    assert E.on_curve(alt) and E.scalar_mult(E.R, alt) == E.IDENTITY


def test_openssl_ed448_public_keys():
    ed448 = pytest.importorskip("cryptography.hazmat.primitives.asymmetric.ed448")
    from cryptography.hazmat.primitives import serialization as S

    rnd = random.Random(1)
    for _ in range(8):
        sk = rnd.randbytes(57)
        pub = ed448.Ed448PrivateKey.from_private_bytes(sk).public_key().public_bytes(S.Encoding.Raw, S.PublicFormat.Raw)
        h = bytearray(hashlib.shake_256(sk).digest(114)[:57])
        h[0] &= 0xFC
        h[55] |= 0x80
        h[56] = 0
        s = int.from_bytes(h, "little")
        assert E.rfc8032_encode(E.scalar_mult(s, E.GENERATOR)) == pub


def test_openssl_x448_variable_base():
    x448 = pytest.importorskip("cryptography.hazmat.primitives.asymmetric.x448")
    rnd = random.Random(2)
    for _ in range(4):
        k = bytearray(rnd.randbytes(56))
        base = E.scalar_mult(rnd.randrange(1, E.R), E.GENERATOR)
        u = E.montgomery_u(base)
        shared = x448.X448PrivateKey.from_private_bytes(bytes(k)).exchange(
            x448.X448PublicKey.from_public_bytes(u.to_bytes(56, "little")))
        k[0] &= 252
        k[55] |= 128
        q = E.scalar_mult(int.from_bytes(k, "little"), base)
        assert E.montgomery_u(q).to_bytes(56, "little") == shared


def test_group_laws():
    rnd = random.Random(3)
    G = E.GENERATOR
    a, b = rnd.randrange(E.R), rnd.randrange(E.R)
    assert E.scalar_mult(a + b, G) == E.point_add(E.scalar_mult(a, G), E.scalar_mult(b, G))
    assert E.scalar_mult(a, E.scalar_mult(b, G)) == E.scalar_mult(b, E.scalar_mult(a, G))
    # unreduced scalar on a point with a 4-torsion component: exact-integer semantics (quirk Q10)
    Pt = E.point_add(G, (1, 0))
    h = rnd.randrange(2**447, 2**448)
    assert E.scalar_mult(h, Pt) == E.point_add(E.scalar_mult(h % E.R, G), E.scalar_mult(h % 4, (1, 0)))


def test_sign_verify_roundtrip_and_rejects():
    """tests/integration_tests.rs:63-81,117-130 restated (D256 and D512)."""
    rnd = random.Random(4)
    for d in (256, 512):
        pw, msg = rnd.randbytes(32), rnd.randbytes(300)
        V = E.keygen(pw, d)
        h, z = E.sign(pw, msg, d)
        assert E.verify(V, msg, h, z, d)
        assert not E.verify(V, msg + b"x", h, z, d)
        assert not E.verify(E.keygen(b"other", d), msg, h, z, d)


def test_ecdhies_roundtrip_and_restore():
    """tests/integration_tests.rs:21-60,268-281 restated."""
    rnd = random.Random(5)
    pw, msg = rnd.randbytes(20), rnd.randbytes(500)
    V = E.keygen(pw, 512)
    ct, t, Z = E.key_encrypt(V, msg, 512, rnd.randbytes(56))
    ok, pt = E.key_decrypt(pw, ct, 512, Z, t)
    assert ok and pt == msg
    ok, buf = E.key_decrypt(b"wrong", ct, 512, Z, t)
    assert not ok and buf == ct


def test_c_oracle_matches_python(oracle):
    rnd = random.Random(6)
    for _ in range(100):
        a, b = rnd.randrange(E.P), rnd.randrange(E.P)
        assert oracle.fe_mul(a.to_bytes(56, "little"), b.to_bytes(56, "little")) == (a * b % E.P).to_bytes(56, "little")
        x, y = rnd.randrange(2**448), rnd.randrange(2**448)
        assert oracle.sc_mul_mod(x.to_bytes(56, "big"), y.to_bytes(56, "big")) == (x * y % E.R).to_bytes(56, "big")
    sc = [rnd.randrange(2**448) for _ in range(6)] + [0, 1, E.R, E.R - 1, 2**448 - 1]
    scb = b"".join(s.to_bytes(56, "big") for s in sc)
    out = oracle.fixed_base_batch(scb)
    for s, row in zip(sc, out):
        assert row.tobytes() == E.point_to_bytes(E.scalar_mult(s, E.GENERATOR))
    pts = [E.scalar_mult(rnd.randrange(E.R), E.GENERATOR) for _ in sc]
    rc, out = oracle.var_base_batch(scb, b"".join(E.point_to_bytes(p) for p in pts))
    assert rc == 0
    for s, p, row in zip(sc, pts, out):
        assert row.tobytes() == E.point_to_bytes(E.scalar_mult(s, p))
    rc, _ = oracle.var_base_batch((5).to_bytes(56, "big"), E.point_to_bytes((5, 7)))
    assert rc == -4


def test_c_oracle_protocol_matches_python(oracle):
    rnd = random.Random(7)
    pws = [rnd.randbytes(rnd.randrange(0, 60)) for _ in range(5)]
    msgs = [rnd.randbytes(rnd.randrange(0, 500)) for _ in range(5)]
    pd, po = cpu.pack(pws)
    md, mo = cpu.pack(msgs)
    for d in (224, 256, 384, 512):
        kg = oracle.keygen_batch(pd, po, d)
        h, z = oracle.sign_batch(pd, po, md, mo, d)
        for i in range(5):
            assert kg[i].tobytes() == E.point_to_bytes(E.keygen(pws[i], d))
            hh, zz = E.sign(pws[i], msgs[i], d)
            assert h[i].tobytes() == hh and z[i].tobytes() == zz
        assert oracle.verify_batch(kg, md, mo, h, z, d).all()
        h2 = h.copy()
        h2[0, 0] ^= 1
        assert list(oracle.verify_batch(kg, md, mo, h2, z, d)) == [0, 1, 1, 1, 1]
    k = rnd.randbytes(56 * 5)
    pub = oracle.keygen_batch(pd, po, 512)
    rc, wx = oracle.ecdh_batch(k, pub)
    assert rc == 0
    for i in range(5):
        kk = int.from_bytes(k[56 * i:56 * i + 56], "big") * 4 % E.R
        assert wx[i].tobytes() == E.fe_to_bytes(E.scalar_mult(kk, E.point_from_bytes(pub[i].tobytes()))[0])

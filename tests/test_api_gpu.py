"""GPU: the Python mirror of the reference's operator surface (capycrypt_b200/api.py) used the way the reference's
own tests use the Rust API (src/sha3/shake_functions.rs:92-203, tests/integration_tests.rs:63-93,117-130)."""
import random

import pytest

from capycrypt_b200.api import Gpu, KeyPair, Message, OperationError, SecParam
from oracle import ref_sha3 as R

pytestmark = pytest.mark.gpu
H = bytes.fromhex


@pytest.fixture(scope="module")
def gpu(engine):
    return Gpu(engine)


def test_sha3_kats_through_message_api(gpu, kat):
    for v in kat["sha3"]:
        m = Message.new(H(v["msg"]))
        gpu.compute_sha3_hash([m], SecParam.try_from(v["d"]))
        assert m.digest.hex() == v["digest"]


def test_tagged_hash_kats(gpu, kat):
    for v in kat["tagged_hash"]:
        m = Message.new(H(v["msg"]))
        gpu.compute_tagged_hash([m], [H(v["pw"])], H(v["s"]), SecParam.try_from(v["d"]))
        assert m.digest.hex() == v["digest"]


def test_try_from_rejects_unsupported():
    with pytest.raises(OperationError) as e:
        SecParam.try_from(128)
    assert e.value.kind == "UnsupportedSecurityParameter"


def test_hashing_twice_mutates_like_the_reference(gpu):
    """Quirk Q5: compute_sha3_hash leaves Message.msg suffixed + padded, so a second call hashes different bytes."""
    rnd = random.Random(1)
    for d in (224, 256, 384, 512):
        for n in (0, 5, 71, 135, 136, 143, 200):
            raw = rnd.randbytes(n)
            m = Message.new(raw)
            gpu.compute_sha3_hash([m], SecParam(d), mutate_like_reference=True)
            buf = bytearray(raw)
            first = R.shake(buf, d)  # the oracle mutates buf exactly as the reference mutates Message.msg
            assert m.digest == first and bytes(m.msg) == bytes(buf)
            gpu.compute_sha3_hash([m], SecParam(d), mutate_like_reference=True)
            assert m.digest == R.shake(buf, d)


def test_sign_verify_like_test_sig_512(gpu):
    rnd = random.Random(2)
    pws = [rnd.randbytes(64) for _ in range(8)]
    keys = gpu.new_keypairs(pws, "test key", SecParam.D512)
    assert all(isinstance(k, KeyPair) and len(k.pub_key) == 112 and k.priv_key == pw for k, pw in zip(keys, pws))
    msgs = [Message.new(rnd.randbytes(5242 + i)) for i in range(8)]
    gpu.sign(msgs, keys, SecParam.D512)
    assert all(m.sig is not None and m.d == SecParam.D512 for m in msgs)
    assert gpu.verify(msgs, [k.pub_key for k in keys]) == [None] * 8
    res = gpu.verify(msgs, [keys[(i + 1) % 8].pub_key for i in range(8)])
    assert all(r is not None and r.kind == "SignatureVerificationFailure" for r in res)
    res = gpu.verify([Message.new(b"x")], [keys[0].pub_key])
    assert res[0].kind == "SignatureNotSet"


def test_scrub_keeps_the_ctx_usable(gpu):
    """capy_gpu_scrub zeroes the device scratch (intermediate secrets); the next calls must work and agree."""
    m = [Message.new(b"abc"), Message.new(b"")]
    gpu.compute_sha3_hash(m, SecParam.D256)
    first = [x.digest for x in m]
    keys = gpu.new_keypairs([b"pw1", b"pw2"], "k", SecParam.D512)
    gpu.engine.scrub()
    m2 = [Message.new(b"abc"), Message.new(b"")]
    gpu.compute_sha3_hash(m2, SecParam.D256)
    assert [x.digest for x in m2] == first
    assert [k.pub_key for k in gpu.new_keypairs([b"pw1", b"pw2"], "k", SecParam.D512)] == [k.pub_key for k in keys]


def test_one_malformed_signature_fails_alone(gpu):
    """A signature whose h is not 56 bytes (or a key that is not 112) fails on its own; the fixed-stride batch the
    C side sees is built from the well-formed items only, so their verdicts do not shift."""
    from capycrypt_b200.api import Signature

    rnd = random.Random(7)
    pws = [rnd.randbytes(16) for _ in range(6)]
    keys = gpu.new_keypairs(pws, "k", SecParam.D256)
    msgs = [Message.new(rnd.randbytes(40 + i)) for i in range(6)]
    gpu.sign(msgs, keys, SecParam.D256)
    pubs = [k.pub_key for k in keys]
    msgs[1].sig = Signature(h=msgs[1].sig.h[:55], z=msgs[1].sig.z)
    msgs[3].sig = Signature(h=msgs[3].sig.h, z=msgs[3].sig.z + b"\0")
    pubs[4] = pubs[4][:111]
    res = gpu.verify(msgs, pubs)
    assert [r is None for r in res] == [True, False, True, False, False, True]
    assert all(r.kind == "SignatureVerificationFailure" for r in res if r is not None)


def test_engine_refuses_short_buffers(engine):
    """The C ABI trusts (pointer, n): the array front end checks n * stride before every call."""
    import numpy as np

    from capycrypt_b200 import pack

    md, mo = pack([b"abc", b"defg"])
    with pytest.raises(ValueError):
        engine.ed448_verify(np.zeros(112, np.uint8), md, mo, np.zeros(112, np.uint8), np.zeros(112, np.uint8), 256)
    with pytest.raises(ValueError):
        engine.ed448_var_base(np.zeros(56, np.uint8), np.zeros(100, np.uint8))
    with pytest.raises(ValueError):
        engine.sha3(md[:5], mo, 256)
    with pytest.raises(ValueError):
        engine.sponge_decrypt(*pack([b"p", b"q"]), np.zeros(1000, np.uint8), 512, md, mo, np.zeros(128, np.uint8), 256)
    with pytest.raises(ValueError):
        engine.ed448_key_decrypt(*pack([b"p", b"q"]), np.zeros(224, np.uint8), md, mo, np.zeros(56, np.uint8), 256)


def test_key_decrypt_with_malformed_nonce_point_fails_alone(gpu):
    rnd = random.Random(8)
    pws = [rnd.randbytes(12) for _ in range(3)]
    keys = gpu.new_keypairs(pws, "k", SecParam.D512)
    plain = [rnd.randbytes(30 + i) for i in range(3)]
    msgs = [Message.new(p) for p in plain]
    gpu.key_encrypt(msgs, [k.pub_key for k in keys], SecParam.D512)
    ct1 = bytes(msgs[1].msg)
    msgs[1].asym_nonce = msgs[1].asym_nonce[:100]
    res = gpu.key_decrypt(msgs, pws)
    assert res[0] is None and res[2] is None and res[1].kind == "KeyDecryptionError"
    assert bytes(msgs[0].msg) == plain[0] and bytes(msgs[2].msg) == plain[2] and bytes(msgs[1].msg) == ct1

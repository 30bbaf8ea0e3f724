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

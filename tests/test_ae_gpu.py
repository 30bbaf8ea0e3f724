"""GPU parity of the authenticated-encryption compositions against the oracle, and the
reference's own encryption tests restated (tests/integration_tests.rs:21-60, 96-114, 238-282).

  sponge AE   SpongeEncryptable::sha3_encrypt / sha3_decrypt   src/sha3/encryptable.rs:29-83   (+ kem/encryptable.rs:51-104)
  ECDHIES     KeyEncryptable::key_encrypt / key_decrypt        src/ecc/encryptable.rs:34-94
Bit-exact: ciphertexts, tags, nonce points, restored buffers and ok flags.
"""
import random

import numpy as np
import pytest

from capycrypt_b200 import _binding as B
from capycrypt_b200.api import Gpu, Message, SecParam
from capycrypt_b200.engine import pack
from oracle import ref_ed448 as E
from oracle import ref_sha3 as R

pytestmark = pytest.mark.gpu

# lengths around every block boundary of the 136- and 168-byte rates, plus empty and long
LENS = [0, 1, 7, 8, 55, 63, 64, 65, 135, 136, 137, 167, 168, 169, 271, 272, 273, 335, 336, 337, 1000, 4096, 5000]


def _split(buf, off):
    return [buf[int(off[i]):int(off[i + 1])].tobytes() for i in range(len(off) - 1)]


@pytest.mark.parametrize("d", [224, 256, 384, 512])
def test_sponge_encrypt_decrypt_matches_oracle(engine, d):
    rnd = random.Random(100 + d)
    msgs = [rnd.randbytes(n) for n in LENS]
    pws = [rnd.randbytes(rnd.choice([0, 1, 16, 64, 200])) for _ in msgs]
    nonces = [rnd.randbytes(512) for _ in msgs]
    pd, po = pack(pws)
    md, mo = pack(msgs)
    ct, tag = engine.sponge_encrypt(pd, po, b"".join(nonces), 512, md, mo, d)
    cts = _split(ct, mo)
    for i, m in enumerate(msgs):
        c_ref, t_ref = R.sha3_encrypt(m, pws[i], d, nonces[i])
        assert cts[i] == c_ref, (d, i, len(m))
        assert tag[i].tobytes() == t_ref, (d, i, len(m))
    # decrypt: every second item gets a wrong password -> failure, buffer = ciphertext
    pws2 = [pw if i % 2 == 0 else pw + b"!" for i, pw in enumerate(pws)]
    pd2, po2 = pack(pws2)
    out, ok = engine.sponge_decrypt(pd2, po2, b"".join(nonces), 512, ct, mo, tag, d)
    outs = _split(out, mo)
    for i, m in enumerate(msgs):
        ok_ref, buf_ref = R.sha3_decrypt(cts[i], pws2[i], d, nonces[i], tag[i].tobytes())
        assert bool(ok[i]) == ok_ref == (i % 2 == 0), (d, i)
        assert outs[i] == buf_ref, (d, i)
        if i % 2 == 0:
            assert outs[i] == m


def test_sponge_kem_variant_uses_kem_customisation(engine):
    """kem/encryptable.rs:51-57: same composition with "KEMKE" / "KEMKA" and the shared secret as the password."""
    rnd = random.Random(7)
    msgs = [rnd.randbytes(n) for n in (0, 33, 136, 1000)]
    secrets = [rnd.randbytes(32) for _ in msgs]
    nonces = [rnd.randbytes(512) for _ in msgs]
    sd, so = pack(secrets)
    md, mo = pack(msgs)
    d = 256
    ct, tag = engine.sponge_encrypt(sd, so, b"".join(nonces), 512, md, mo, d, variant=B.AE_KEM)
    cts = _split(ct, mo)
    for i, m in enumerate(msgs):
        ke_ka = R.kmac_xof(nonces[i] + secrets[i], b"", 1024, b"S", d)
        ke, ka = ke_ka[:64], ke_ka[64:]
        assert tag[i].tobytes() == R.kmac_xof(ka, m, 512, b"KEMKA", d)
        ks = R.kmac_xof(ke, b"", len(m) * 8, b"KEMKE", d)
        assert cts[i] == bytes(a ^ b for a, b in zip(m, ks))
    out, ok = engine.sponge_decrypt(sd, so, b"".join(nonces), 512, ct, mo, tag, d, variant=B.AE_KEM)
    assert ok.tolist() == [1] * len(msgs) and _split(out, mo) == msgs
    # the SHA3 variant must not open a KEM ciphertext
    out, ok = engine.sponge_decrypt(sd, so, b"".join(nonces), 512, ct, mo, tag, d, variant=B.AE_SHA3)
    assert ok.tolist() == [0] * len(msgs) and _split(out, mo) == cts


def test_sponge_ae_other_nonce_lengths_and_bad_args(engine):
    rnd = random.Random(8)
    msgs = [rnd.randbytes(100) for _ in range(3)]
    pws = [b"pw"] * 3
    for nl in (0, 16, 136):
        nonces = [rnd.randbytes(nl) for _ in msgs]
        pd, po = pack(pws)
        md, mo = pack(msgs)
        ct, tag = engine.sponge_encrypt(pd, po, b"".join(nonces), nl, md, mo, 512)
        for i, m in enumerate(msgs):
            c_ref, t_ref = R.sha3_encrypt(m, pws[i], 512, nonces[i])
            assert _split(ct, mo)[i] == c_ref and tag[i].tobytes() == t_ref
    with pytest.raises(B.CapyError) as e:
        engine.sponge_encrypt(pd, po, b"", 0, md, mo, 128)
    assert e.value.status == B.ERR_BAD_SECPARAM
    with pytest.raises(B.CapyError) as e:
        engine.sponge_encrypt(pd, po, b"", 0, md, mo, 256, variant=9)
    assert e.value.status == B.ERR_BAD_ARG
    # empty batch
    ct, tag = engine.sponge_encrypt(np.zeros(0, np.uint8), np.zeros(1, np.uint64), b"", 512, np.zeros(0, np.uint8),
                                    np.zeros(1, np.uint64), 256)
    assert len(ct) == 0 and tag.shape == (0, 64)


def test_sponge_roundtrip_large_ragged_batch(engine):
    """Size-independent property at a larger size: decrypt(encrypt(m)) == m for 4096 ragged messages, tags differ
    per message, a flipped ciphertext bit is rejected and handed back unchanged."""
    rng = np.random.default_rng(9)
    n = 4096
    lens = rng.integers(0, 3000, n)
    off = np.zeros(n + 1, np.uint64)
    off[1:] = np.cumsum(lens)
    md = rng.integers(0, 256, int(off[-1]), dtype=np.uint8)
    pd = rng.integers(0, 256, 16 * n, dtype=np.uint8)
    po = np.arange(n + 1, dtype=np.uint64) * 16
    nonces = rng.integers(0, 256, 512 * n, dtype=np.uint8)
    ct, tag = engine.sponge_encrypt(pd, po, nonces, 512, md, off, 256)
    out, ok = engine.sponge_decrypt(pd, po, nonces, 512, ct, off, tag, 256)
    assert ok.all() and np.array_equal(out, md)
    assert len({t.tobytes() for t in tag}) == n
    # spot-check 16 items against the oracle
    for i in rng.integers(0, n, 16):
        m = md[int(off[i]):int(off[i + 1])].tobytes()
        c_ref, t_ref = R.sha3_encrypt(m, pd[16 * i:16 * i + 16].tobytes(), 256, nonces[512 * i:512 * i + 512].tobytes())
        assert ct[int(off[i]):int(off[i + 1])].tobytes() == c_ref and tag[i].tobytes() == t_ref
    bad = ct.copy()
    nz = np.nonzero(lens)[0]
    for i in nz[:50]:
        bad[int(off[i])] ^= 1
    out, ok = engine.sponge_decrypt(pd, po, nonces, 512, bad, off, tag, 256)
    assert not ok[nz[:50]].any() and ok.sum() == n - 50
    for i in nz[:50]:
        assert np.array_equal(out[int(off[i]):int(off[i + 1])], bad[int(off[i]):int(off[i + 1])])


@pytest.mark.parametrize("d", [256, 512])
def test_key_encrypt_decrypt_matches_oracle(engine, d):
    rnd = random.Random(200 + d)
    lens = [0, 1, 55, 56, 57, 135, 136, 137, 168, 500, 2000]
    msgs = [rnd.randbytes(n) for n in lens]
    pws = [rnd.randbytes(rnd.choice([0, 8, 32, 64])) for _ in msgs]
    k_rand = [rnd.randbytes(56) for _ in msgs]
    k_rand[0] = b"\xff" * 56  # unreduced scalar input
    k_rand[1] = bytes(56)     # k = 0: W = Z = identity
    pubs = [E.keygen(pw, d) for pw in pws]
    pub_b = b"".join(E.point_to_bytes(p) for p in pubs)
    md, mo = pack(msgs)
    rc, ct, tag, z = engine.ed448_key_encrypt(pub_b, b"".join(k_rand), md, mo, d)
    assert rc == 0
    cts = _split(ct, mo)
    for i, m in enumerate(msgs):
        c_ref, t_ref, z_ref = E.key_encrypt(pubs[i], m, d, k_rand[i])
        assert cts[i] == c_ref, (d, i)
        assert tag[i].tobytes() == t_ref, (d, i)
        assert z[i].tobytes() == E.point_to_bytes(z_ref), (d, i)
    pws2 = [pw if i % 3 else pw + b"x" for i, pw in enumerate(pws)]
    pd2, po2 = pack(pws2)
    rc, out, ok = engine.ed448_key_decrypt(pd2, po2, z, ct, mo, tag, d)
    assert rc == 0
    outs = _split(out, mo)
    for i, m in enumerate(msgs):
        ok_ref, buf_ref = E.key_decrypt(pws2[i], cts[i], d, E.point_from_bytes(z[i].tobytes()), tag[i].tobytes())
        assert bool(ok[i]) == ok_ref, (d, i)
        assert outs[i] == buf_ref, (d, i)


def test_key_encrypt_rejects_off_curve_points(engine):
    pub = bytearray(E.point_to_bytes(E.GENERATOR) * 2)
    pub[112] ^= 1  # second key is off-curve
    md, mo = pack([b"abc", b"def"])
    rc, ct, tag, z = engine.ed448_key_encrypt(bytes(pub), b"\x01" * 112, md, mo, 256)
    assert rc == B.ERR_BAD_POINT
    c_ref, t_ref, _ = E.key_encrypt(E.GENERATOR, b"abc", 256, b"\x01" * 56)
    assert ct[:3].tobytes() == c_ref and tag[0].tobytes() == t_ref
    # decrypt with an off-curve nonce point fails for that item and keeps the ciphertext
    rc, out, ok = engine.ed448_key_decrypt(*pack([b"pw", b"pw"]), bytes(pub), ct, mo, tag, 256)
    assert rc == B.ERR_BAD_POINT and ok[1] == 0 and out[3:].tobytes() == ct[3:].tobytes()


# ---- the reference's own tests, restated on the batch API -----------------------------------------------------
@pytest.fixture(scope="module")
def gpu(engine):
    return Gpu(engine)


def test_symmetric_encryptable(gpu):
    """tests/integration_tests.rs:96-114 (SHA3 half), batched, for every SecParam."""
    rnd = random.Random(11)
    for d in (SecParam.D224, SecParam.D256, SecParam.D384, SecParam.D512):
        raw = [rnd.randbytes(5242) for _ in range(4)]
        pws = [rnd.randbytes(64) for _ in raw]
        msgs = [Message.new(r) for r in raw]
        gpu.sha3_encrypt(msgs, pws, d)
        assert all(bytes(m.msg) != r and len(m.msg) == len(r) and len(m.sym_nonce) == 512 and len(m.digest) == 64
                   for m, r in zip(msgs, raw))
        assert gpu.sha3_decrypt(msgs, pws) == [None] * 4
        assert [bytes(m.msg) for m in msgs] == raw


def test_sha3_decrypt_handling_bad_input(gpu):
    """tests/integration_tests.rs:250-262: a failed decryption leaves the encrypted text unchanged."""
    rnd = random.Random(12)
    msgs = [Message.new(rnd.randbytes(523)) for _ in range(3)]
    gpu.sha3_encrypt(msgs, [rnd.randbytes(64) for _ in msgs], SecParam.D512)
    enc = [bytes(m.msg) for m in msgs]
    res = gpu.sha3_decrypt(msgs, [rnd.randbytes(64) for _ in msgs])
    assert all(r is not None and r.kind == "SHA3DecryptionFailure" for r in res)
    assert [bytes(m.msg) for m in msgs] == enc
    m = Message.new(b"abc")
    assert gpu.sha3_decrypt([m], [b""])[0].kind == "SecurityParameterNotSet"
    m.d = SecParam.D256
    assert gpu.sha3_decrypt([m], [b""])[0].kind == "SymNonceNotSet"


@pytest.mark.parametrize("d", [SecParam.D256, SecParam.D512])
def test_key_gen_enc_dec(gpu, d):
    """tests/integration_tests.rs:21-60."""
    rnd = random.Random(13)
    pws = [rnd.randbytes(32) for _ in range(4)]
    keys = gpu.new_keypairs(pws, "test key", d)
    raw = [rnd.randbytes(5242) for _ in keys]
    msgs = [Message.new(r) for r in raw]
    gpu.key_encrypt(msgs, [k.pub_key for k in keys], d)
    assert all(bytes(m.msg) != r and len(m.asym_nonce) == 112 and len(m.digest) == 56 for m, r in zip(msgs, raw))
    assert gpu.key_decrypt(msgs, [k.priv_key for k in keys]) == [None] * 4
    assert [bytes(m.msg) for m in msgs] == raw


def test_key_decrypt_handling_bad_input(gpu):
    """tests/integration_tests.rs:268-281."""
    rnd = random.Random(14)
    k1 = gpu.new_keypairs([rnd.randbytes(32) for _ in range(3)], "test key", SecParam.D512)
    k2 = gpu.new_keypairs([rnd.randbytes(32) for _ in range(3)], "test key", SecParam.D512)
    msgs = [Message.new(rnd.randbytes(125)) for _ in k1]
    gpu.key_encrypt(msgs, [k.pub_key for k in k1], SecParam.D512)
    enc = [bytes(m.msg) for m in msgs]
    res = gpu.key_decrypt(msgs, [k.priv_key for k in k2])
    assert all(r is not None and r.kind == "KeyDecryptionError" for r in res)
    assert [bytes(m.msg) for m in msgs] == enc
    assert gpu.key_decrypt([Message.new(b"x")], [b""])[0].kind == "SymNonceNotSet"


@pytest.mark.parametrize("d", [224, 384])
def test_key_encrypt_large_batch_against_c_oracle(engine, oracle, d):
    """ECDHIES for a few hundred items (the two security parameters the small test above leaves out), expected values
    composed from the C oracle: W.x from its ECDH, (ke || ka) = KMACXOF(W.x, "", 896, "PK"), t = KMACXOF(ka, m, 448, "PKA"),
    c = KMACXOF(ke, "", |m|, "PKE") xor m  (ecc/encryptable.rs:36-45)."""
    rng = np.random.default_rng(300 + d)
    n = 200
    msgs = [bytes(rng.integers(0, 256, size=int(k), dtype=np.uint8)) for k in rng.integers(0, 600, size=n)]
    pws = [bytes(rng.integers(0, 256, size=int(k), dtype=np.uint8)) for k in rng.integers(0, 40, size=n)]
    pd, po = pack(pws)
    md, mo = pack(msgs)
    pub = oracle.keygen_batch(pd, po, d, threads=0)
    k_rand = rng.integers(0, 256, size=n * 56, dtype=np.uint8)
    rc, ct, tag, z = engine.ed448_key_encrypt(pub, k_rand, md, mo, d)
    assert rc == 0
    rc2, wx = oracle.ecdh_batch(k_rand, pub, threads=0)
    assert rc2 == 0
    none, empty = np.zeros(0, np.uint8), np.zeros(n + 1, np.uint64)
    ke_ka = oracle.kmac_xof_batch(wx.reshape(-1), np.arange(n + 1, dtype=np.uint64) * 56, none, empty, 896, b"PK", d, threads=0)
    ka, kao = pack([r[56:].tobytes() for r in ke_ka])
    assert np.array_equal(tag, oracle.kmac_xof_batch(ka, kao, md, mo, 448, b"PKA", d, threads=0))
    for i in rng.choice(n, 40, replace=False):
        if not msgs[i]:
            continue
        ks = oracle.kmac_xof_batch(ke_ka[i, :56], np.array([0, 56], np.uint64), none, np.zeros(2, np.uint64), 8 * len(msgs[i]), b"PKE",
                                   d, threads=1)[0]
        assert np.array_equal(ct[int(mo[i]):int(mo[i + 1])], ks ^ np.frombuffer(msgs[i], np.uint8)), (d, i)
    # Z = [4 k mod r] G : decrypting with the right passwords must open everything
    rc, out, ok = engine.ed448_key_decrypt(pd, po, z, ct, mo, tag, d)
    assert rc == 0 and ok.all() and np.array_equal(out, md)

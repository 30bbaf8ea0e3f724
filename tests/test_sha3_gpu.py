"""GPU parity tests for the Keccak sponge path: CUDA engine (through the C ABI) vs the CPU oracle
on identical seeded inputs; bit-exact (byte work).  Mirrors the reference's own tests
(src/sha3/shake_functions.rs:92-288) and adds the quirk lengths of SURVEY.md App. A."""
import hashlib
import random

import numpy as np
import pytest

from capycrypt_b200 import pack
from capycrypt_b200 import _binding as B

pytestmark = pytest.mark.gpu
H = bytes.fromhex
DS = (224, 256, 384, 512)


# ---- the reference's own known-answer tests, through the C ABI ------------------------------------
def test_kat_sha3(engine, kat):
    for v in kat["sha3"]:
        data, off = pack([H(v["msg"])])
        assert engine.sha3(data, off, v["d"])[0].tobytes().hex() == v["digest"], v["src"]


def test_kat_tagged_hash(engine, kat):
    for v in kat["tagged_hash"]:
        kd, ko = pack([H(v["pw"])])
        xd, xo = pack([H(v["msg"])])
        out = engine.kmac_xof(kd, ko, xd, xo, v["d"], H(v["s"]), v["d"])
        assert out[0].tobytes().hex() == v["digest"]


def test_kat_cshake(engine, kat):
    for v in kat["cshake"]:
        xd, xo = pack([H(v["x"])])
        out = engine.cshake(xd, xo, v["l"], H(v["n"]), H(v["s"]), v["d"])
        assert out[0].tobytes().hex() == v["out"]


def test_kat_kmac_xof(engine, kat):
    for v in kat["kmac_xof"]:
        kd, ko = pack([H(v["k"])])
        xd, xo = pack([H(v["x"])])
        out = engine.kmac_xof(kd, ko, xd, xo, v["l"], H(v["s"]), v["d"])
        assert out[0].tobytes().hex() == v["out"]


# ---- error behaviour --------------------------------------------------------------------------------
def test_unsupported_security_parameter(engine):
    data, off = pack([b"abc"])
    for d in (0, 128, 255, 1024):
        with pytest.raises(B.CapyError) as e:
            engine.sha3(data, off, d)
        assert e.value.status == B.ERR_BAD_SECPARAM
        with pytest.raises(B.CapyError) as e:
            engine.kmac_xof(data, off, data, off, 256, b"", d)
        assert e.value.status == B.ERR_BAD_SECPARAM


def test_empty_batch(engine):
    out = engine.sha3(np.zeros(0, np.uint8), np.zeros(1, np.uint64), 256)
    assert out.shape == (0, 32)


# ---- SHA3-d sweeps ------------------------------------------------------------------------------------
@pytest.mark.parametrize("d", DS)
def test_sha3_every_length_0_to_600(engine, oracle, d):
    """Every length 0..600 in ONE ragged batch (covers empty input, every pad position, and the
    quirk lengths len % r == r-1 and len % 136 == 135 for every rate)."""
    rnd = random.Random(100 + d)
    msgs = [rnd.randbytes(n) for n in range(0, 601)]
    data, off = pack(msgs)
    got = engine.sha3(data, off, d)
    want = oracle.sha3_batch(data, off, d)
    assert np.array_equal(got, want), np.nonzero((got != want).any(axis=1))[0][:10]


def test_sha3_quirk_lengths_differ_from_fips(engine):
    """Quirk Q1/Q2: at these lengths the reference (and therefore the engine) is NOT FIPS 202."""
    for n in (71, 135, 143, 215):
        data, off = pack([bytes(n)])
        got = engine.sha3(data, off, 512)[0].tobytes()
        assert got != hashlib.sha3_512(bytes(n)).digest()
    for n in (0, 1, 70, 72, 134, 136):
        data, off = pack([bytes(n)])
        assert engine.sha3(data, off, 512)[0].tobytes() == hashlib.sha3_512(bytes(n)).digest()


def test_sha3_256_matches_hashlib(engine):
    rnd = random.Random(5)
    msgs = [rnd.randbytes(rnd.randrange(0, 5000)) for _ in range(300)]
    data, off = pack(msgs)
    got = engine.sha3(data, off, 256)
    for m, g in zip(msgs, got):
        assert g.tobytes() == hashlib.sha3_256(m).digest()


@pytest.mark.parametrize("d", DS)
def test_sha3_ragged_long(engine, oracle, d):
    rnd = random.Random(200 + d)
    lens = [rnd.randrange(0, 20000) for _ in range(500)] + [0, 1, 100000, 65536, 72 * 100 - 1, 136 * 50 - 1]
    rnd.shuffle(lens)
    msgs = [rnd.randbytes(n) for n in lens]
    data, off = pack(msgs)
    assert np.array_equal(engine.sha3(data, off, d), oracle.sha3_batch(data, off, d))


@pytest.mark.parametrize("d", DS)
@pytest.mark.parametrize("msg_len,stride", [(64, 64), (32, 32), (32, 48), (64, 80), (0, 8), (1, 8), (63, 64), (71, 72), (135, 136), (136, 136),
                                             (143, 144), (200, 200), (1000, 1000), (4096, 4096), (100, 128),
                                             (64, 65), (33, 33),
                                             # every single-block length the compile-time specialised kernel takes
                                             (16, 16), (48, 48), (80, 80), (96, 112), (112, 112), (128, 128), (128, 144)])
def test_sha3_fixed(engine, oracle, d, msg_len, stride):
    """Uniform-length entry point (cfg-1 shape), aligned strides (fast kernel) and odd strides."""
    n = 777
    rnd = np.random.default_rng(d + msg_len)
    buf = rnd.integers(0, 256, size=n * stride, dtype=np.uint8)
    got = engine.sha3_fixed(buf, msg_len, stride, n, d)
    off = np.zeros(n + 1, np.uint64)
    msgs = [buf[i * stride:i * stride + msg_len].tobytes() for i in range(n)]
    data, off = pack(msgs)
    assert np.array_equal(got, oracle.sha3_batch(data, off, d))


def test_sha3_cfg1_full_size(engine, oracle):
    """BASELINE config 1 at full size: 2^20 x 64 B, SHA3-256.  A seeded sample goes to the oracle;
    the whole batch is checked through a size-independent property: permuting the batch permutes
    the digests (every item is independent of its neighbours and of its position)."""
    n = 1 << 20
    rnd = np.random.default_rng(1)
    buf = rnd.integers(0, 256, size=n * 64, dtype=np.uint8)
    got = engine.sha3_fixed(buf, 64, 64, n, 256)
    idx = rnd.choice(n, size=4096, replace=False)
    sample = np.ascontiguousarray(buf.reshape(n, 64)[idx]).reshape(-1)
    off = np.arange(4097, dtype=np.uint64) * 64
    assert np.array_equal(got[idx], oracle.sha3_batch(sample, off, 256))
    perm = rnd.permutation(n)
    got2 = engine.sha3_fixed(np.ascontiguousarray(buf.reshape(n, 64)[perm]).reshape(-1), 64, 64, n, 256)
    assert np.array_equal(got2, got[perm])
    # the offsets form gives the same digests as the fixed form
    got3 = engine.sha3(buf, np.arange(n + 1, dtype=np.uint64) * 64, 256)
    assert np.array_equal(got3, got)


# ---- cSHAKE / KMACXOF sweeps -----------------------------------------------------------------------------
@pytest.mark.parametrize("d", DS)
def test_kmac_xof_sweep(engine, oracle, d):
    rnd = random.Random(300 + d)
    klens = [0, 1, 32, 56, 130, 131, 132, 135, 163, 164, 167, 168, 267, 539] + [rnd.randrange(0, 400) for _ in range(40)]
    for l_bits in (8, 64, 448, 512, 1024, 1088, 1344, 1352, 4096, 8 * 700):
        keys = [rnd.randbytes(rnd.choice(klens)) for _ in range(96)]
        msgs = [rnd.randbytes(rnd.choice([0, 1, 7, 8, 100, 133, 134, 135, 136, 165, 166, 167, 168, rnd.randrange(0, 3000)]))
                for _ in range(96)]
        s = rnd.randbytes(rnd.choice([0, 1, 2, 21, 100, 140]))
        kd, ko = pack(keys)
        xd, xo = pack(msgs)
        got = engine.kmac_xof(kd, ko, xd, xo, l_bits, s, d)
        want = oracle.kmac_xof_batch(kd, ko, xd, xo, l_bits, s, d)
        assert np.array_equal(got, want), (d, l_bits, np.nonzero((got != want).any(axis=1))[0][:10])


@pytest.mark.parametrize("d", DS)
def test_cshake_sweep_including_q4(engine, oracle, d):
    rnd = random.Random(400 + d)
    for fn, cs in ((b"", b"Email Signature"), (b"KMAC", b""), (b"abc", b"xyz" * 50), (b"", b"")):  # last = quirk Q4
        msgs = [rnd.randbytes(n) for n in list(range(0, 360)) + [rnd.randrange(360, 5000) for _ in range(40)]]
        xd, xo = pack(msgs)
        for l_bits in (256, 512, 2048):
            got = engine.cshake(xd, xo, l_bits, fn, cs, d)
            want = oracle.cshake_batch(xd, xo, l_bits, fn, cs, d)
            assert np.array_equal(got, want), (d, fn, cs, l_bits, np.nonzero((got != want).any(axis=1))[0][:10])


def test_kmac_variable_output_lengths(engine, oracle):
    """Variable-length squeeze per item (keystream shape of sha3/encryptable.rs:41)."""
    rnd = random.Random(9)
    n = 64
    keys = [rnd.randbytes(64) for _ in range(n)]
    msgs = [b"" for _ in range(n)]
    out_lens = [rnd.choice([0, 1, 8, 135, 136, 137, 1000, 5000]) for _ in range(n)]
    out_off = np.zeros(n + 1, np.uint64)
    out_off[1:] = np.cumsum(out_lens)
    kd, ko = pack(keys)
    xd, xo = pack(msgs)
    got = engine.kmac_xof(kd, ko, xd, xo, 0, b"SKE", 512, out_off=out_off)
    for i in range(n):
        want = oracle.kmac_xof(keys[i], b"", out_lens[i] * 8, b"SKE", 512)
        assert got[int(out_off[i]):int(out_off[i + 1])].tobytes() == want, i


def test_kmac_cfg2_shape(engine, oracle):
    """BASELINE config 2 shape (4 KB messages, 32-byte keys, D512) at a size the oracle finishes
    in seconds, all four output lengths of the config."""
    n = 2048
    rnd = np.random.default_rng(2)
    xd = rnd.integers(0, 256, size=n * 4096, dtype=np.uint8)
    kd = rnd.integers(0, 256, size=n * 32, dtype=np.uint8)
    xo = np.arange(n + 1, dtype=np.uint64) * 4096
    ko = np.arange(n + 1, dtype=np.uint64) * 32
    for l_bits in (256, 512, 4096, 32768):
        for s in (b"My Tagged Application", b""):
            got = engine.kmac_xof(kd, ko, xd, xo, l_bits, s, 512)
            want = oracle.kmac_xof_batch(kd, ko, xd, xo, l_bits, s, 512, threads=0)
            assert np.array_equal(got, want), (l_bits, s)


def test_unaligned_message_starts(engine, oracle):
    """Packed batches whose items start at every byte phase (exercises the funnel-shift load path)."""
    rnd = random.Random(11)
    msgs = [rnd.randbytes(rnd.choice([137, 273, 409, 1001, 2049, 555])) for _ in range(257)]
    data, off = pack(msgs)
    for d in (256, 512):
        assert np.array_equal(engine.sha3(data, off, d), oracle.sha3_batch(data, off, d))
    kd, ko = pack([rnd.randbytes(33) for _ in msgs])
    got = engine.kmac_xof(kd, ko, data, off, 512, b"T", 512)
    assert np.array_equal(got, oracle.kmac_xof_batch(kd, ko, data, off, 512, b"T", 512))


# ---- device-pointer API + FIPS SHAKE extra ------------------------------------------------------------------
def test_device_pointer_api_and_shake_extra(engine, oracle):
    import torch

    rnd = random.Random(12)
    msgs = [rnd.randbytes(rnd.randrange(0, 700)) for _ in range(500)]
    data, off = pack(msgs)
    t_data = torch.from_numpy(np.concatenate([data, np.zeros(8, np.uint8)])).cuda()
    t_off = torch.from_numpy(off.astype(np.int64)).cuda()
    t_out = torch.zeros(len(msgs) * 64, dtype=torch.uint8, device="cuda")
    engine.sha3_dev(t_data, t_off, 512, t_out)
    torch.cuda.synchronize()
    assert np.array_equal(t_out.cpu().numpy().reshape(-1, 64), oracle.sha3_batch(data, off, 512))
    for bits, ref in ((256, hashlib.shake_256), (128, hashlib.shake_128)):
        t_out = torch.zeros(len(msgs) * 200, dtype=torch.uint8, device="cuda")
        engine.fips_shake_dev(t_data, t_off, bits, 200, t_out)
        torch.cuda.synchronize()
        got = t_out.cpu().numpy().reshape(-1, 200)
        for m, g in zip(msgs, got):
            assert g.tobytes() == ref(m).digest(200)


# ---- mixed-size batches (BASELINE config 5 shape): longest-first scheduling must not change any digest ----
def test_mixed_size_sha3_512_vs_oracle(engine, oracle):
    """Log-uniform lengths in [64 B, 1 MiB] incl. the quirk lengths len % 72 == 71 and len % 136 == 135;
    every digest against the oracle; then the same batch with CAPY_FLAG_NO_SORT and permuted."""
    import torch

    rnd = np.random.default_rng(5)
    lens = np.exp(rnd.uniform(np.log(64), np.log(1 << 20), size=600)).astype(np.int64)
    lens[:8] = [71, 143, 135, 271, 72 * 1000 + 71, 136 * 500 + 135, 1 << 20, 64]
    off = np.zeros(len(lens) + 1, np.uint64)
    off[1:] = np.cumsum(lens)
    data = rnd.integers(0, 256, size=int(off[-1]), dtype=np.uint8)
    want = oracle.sha3_batch(data, off, 512, threads=0)
    got = engine.sha3(data, off, 512)
    assert np.array_equal(got, want), np.nonzero((got != want).any(axis=1))[0][:10]
    t_data = torch.from_numpy(np.concatenate([data, np.zeros(8, np.uint8)])).cuda()
    t_off = torch.from_numpy(off.astype(np.int64)).cuda()
    t_out = torch.zeros(len(lens) * 64, dtype=torch.uint8, device="cuda")
    engine._check(engine.lib.capy_sha3_batch_dev(engine._ctx, 0, engine._stream(), 512, t_data.data_ptr(), t_off.data_ptr(),
                                                 len(lens), t_out.data_ptr(), 1))  # CAPY_FLAG_NO_SORT
    torch.cuda.synchronize()
    assert np.array_equal(t_out.cpu().numpy().reshape(-1, 64), want)


def test_mixed_size_many_short_and_few_long(engine, oracle):
    rnd = np.random.default_rng(6)
    lens = np.concatenate([rnd.integers(0, 300, size=5000), rnd.integers(200000, 400000, size=5), [0, 0, 1]])
    rnd.shuffle(lens)
    off = np.zeros(len(lens) + 1, np.uint64)
    off[1:] = np.cumsum(lens)
    data = rnd.integers(0, 256, size=int(off[-1]), dtype=np.uint8)
    for d in (256, 512):
        assert np.array_equal(engine.sha3(data, off, d), oracle.sha3_batch(data, off, d, threads=0))

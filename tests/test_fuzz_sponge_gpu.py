"""GPU: randomised differential test of the sponge entry points against the C oracle (oracle/ref_cpu.c, itself pinned
by the reference's KATs): random security parameter, customisation strings, key / message / output lengths drawn
around every block boundary (rates 72..172, the 172-byte quirk of D224 included), ragged batches that mix empty, short
and long items so that all three launch tiers and both load paths (aligned / byte-phase) are hit.  Bit-exact."""
import numpy as np
import pytest

from capycrypt_b200 import pack

pytestmark = pytest.mark.gpu
RATES = [72, 104, 136, 144, 152, 168, 172]


def _len(rng, long_ok=True):
    kind = rng.integers(0, 10)
    if kind < 3:
        return int(rng.integers(0, 40))
    if kind < 8:
        r = RATES[rng.integers(0, len(RATES))]
        return max(0, int(r * rng.integers(0, 6) + rng.integers(-3, 4)))
    if kind == 8 or not long_ok:
        return int(rng.integers(200, 5000))
    return int(rng.integers(20_000, 120_000))


def _items(rng, n, long_ok=True):
    return [bytes(rng.integers(0, 256, size=_len(rng, long_ok), dtype=np.uint8)) for _ in range(n)]


@pytest.mark.parametrize("seed", range(6))
def test_fuzz_sha3(engine, oracle, seed):
    rng = np.random.default_rng(1000 + seed)
    for d in (224, 256, 384, 512):
        n = int(rng.integers(1, 400))
        data, off = pack(_items(rng, n))
        got = engine.sha3(data, off, d)
        want = oracle.sha3_batch(data, off, d, threads=0)
        assert np.array_equal(got, want), (seed, d, np.nonzero((got != want).any(axis=1))[0][:5])


@pytest.mark.parametrize("seed", range(6))
def test_fuzz_kmac_and_cshake(engine, oracle, seed):
    rng = np.random.default_rng(2000 + seed)
    for d in (224, 256, 384, 512):
        n = int(rng.integers(1, 200))
        data, off = pack(_items(rng, n))
        keys, koff = pack(_items(rng, n, long_ok=False))
        custom = bytes(rng.integers(0, 256, size=int(rng.integers(0, 200)), dtype=np.uint8))
        out_bits = 8 * int(rng.choice([1, 28, 56, 64, 135, 136, 137, 168, 172, 400, 1000]))
        got = engine.kmac_xof(keys, koff, data, off, out_bits, custom, d)
        want = oracle.kmac_xof_batch(keys, koff, data, off, out_bits, custom, d, threads=0)
        assert np.array_equal(got, want), ("kmac", seed, d, out_bits, np.nonzero((got != want).any(axis=1))[0][:5])
        fn = bytes(rng.integers(0, 256, size=int(rng.integers(0, 40)), dtype=np.uint8))
        got = engine.cshake(data, off, out_bits, fn, custom, d)
        want = oracle.cshake_batch(data, off, out_bits, fn, custom, d, threads=0)
        assert np.array_equal(got, want), ("cshake", seed, d, out_bits, np.nonzero((got != want).any(axis=1))[0][:5])


def test_fuzz_kmac_variable_output_lengths(engine, oracle):
    """keystream shape: per-item output lengths (out_off), some of them long enough for the fast tiers"""
    rng = np.random.default_rng(3000)
    for d in (256, 512):
        n = 60
        data, off = pack(_items(rng, n, long_ok=False))
        keys, koff = pack(_items(rng, n, long_ok=False))
        outs = [_len(rng) for _ in range(n)]
        out_off = np.zeros(n + 1, np.uint64)
        out_off[1:] = np.cumsum(outs)
        got = engine.kmac_xof(keys, koff, data, off, 0, b"SKE", d, out_off=out_off)
        for i in range(n):
            want = oracle.kmac_xof_batch(keys[int(koff[i]):int(koff[i + 1])], np.array([0, koff[i + 1] - koff[i]], np.uint64),
                                         data[int(off[i]):int(off[i + 1])], np.array([0, off[i + 1] - off[i]], np.uint64),
                                         8 * outs[i], b"SKE", d, threads=1)[0] if outs[i] else np.zeros(0, np.uint8)
            assert np.array_equal(got[int(out_off[i]):int(out_off[i + 1])], want), (d, i, outs[i])


@pytest.mark.parametrize("d", [224, 256, 384, 512])
def test_fuzz_sponge_ae_against_kmac_oracle(engine, oracle, d):
    """sha3_encrypt / sha3_decrypt (sha3/encryptable.rs:29-83) for a few hundred ragged items, the expected values composed
    from the C oracle's KMACXOF: (ke || ka) = KMACXOF(z || pw, "", 1024, "S"), t = KMACXOF(ka, m, 512, "SKA"),
    c = KMACXOF(ke, "", |m|, "SKE") xor m."""
    rng = np.random.default_rng(4000 + d)
    n = 300
    msgs = _items(rng, n)
    pws = _items(rng, n, long_ok=False)
    nonces = [bytes(rng.integers(0, 256, size=512, dtype=np.uint8)) for _ in range(n)]
    md, mo = pack(msgs)
    pd, po = pack(pws)
    ct, tag = engine.sponge_encrypt(pd, po, b"".join(nonces), 512, md, mo, d)
    kd, ko = pack([z + p for z, p in zip(nonces, pws)])
    empty = np.zeros(n + 1, np.uint64)
    none = np.zeros(0, np.uint8)
    ke_ka = oracle.kmac_xof_batch(kd, ko, none, empty, 1024, b"S", d, threads=0)
    ka, kao = pack([r[64:].tobytes() for r in ke_ka])
    want_tag = oracle.kmac_xof_batch(ka, kao, md, mo, 512, b"SKA", d, threads=0)
    assert np.array_equal(tag, want_tag), np.nonzero((tag != want_tag).any(axis=1))[0][:5]
    for i in rng.choice(n, 60, replace=False):  # keystream per item (the oracle squeezes one length per call)
        m = msgs[i]
        if not m:
            continue
        ks = oracle.kmac_xof_batch(ke_ka[i, :64], np.array([0, 64], np.uint64), none, np.zeros(2, np.uint64), 8 * len(m), b"SKE", d,
                                   threads=1)[0]
        assert np.array_equal(ct[int(mo[i]):int(mo[i + 1])], ks ^ np.frombuffer(m, np.uint8)), (d, i, len(m))
    out, ok = engine.sponge_decrypt(pd, po, b"".join(nonces), 512, ct, mo, tag, d)
    assert ok.all() and np.array_equal(out, md)

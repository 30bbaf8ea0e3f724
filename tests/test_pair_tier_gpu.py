"""GPU: the fast tiers of chain-bound ragged sponge batches (csrc/keccak_pair.cuh, sponge_tiered_kernel).

A sponge is sequential per message, so a batch whose longest chain outlasts the work-bound time runs its longest
items with a whole warp per message (one lane per thread) or with two adjacent threads per message (low / high 32
bits of every lane); the small batches below land in the warp tier, the larger ones spread over all three tiers.  Every digest, tag and
keystream byte must be identical to the thread-per-item path and to the oracle (sha3/sponge.rs:10-34,
sha3/shake_functions.rs:24-89), including the reference's padding quirks at the block boundaries."""
import random

import numpy as np
import pytest

from capycrypt_b200.engine import pack
from oracle import ref_sha3 as R

pytestmark = pytest.mark.gpu


def _ragged(rnd, lens):
    off = np.zeros(len(lens) + 1, np.uint64)
    off[1:] = np.cumsum(lens)
    data = rnd.integers(0, 256, size=int(off[-1]), dtype=np.uint8)
    return data, off


@pytest.mark.parametrize("d", [224, 256, 384, 512])
def test_chain_bound_sha3_batches(engine, oracle, d):
    """few long + many short messages: the long ones (incl. block-boundary and quirk lengths) run as thread pairs"""
    rnd = np.random.default_rng(40 + d)
    r = {224: 144, 256: 136, 384: 104, 512: 72}[d]
    long_lens = [200_000, 150_001, 120_000 + 3, r * 1500, r * 1500 - 1, r * 1500 + 1, 136 * 900 + 135, r * 1200 + r - 1,
                 90_007, 90_006, 90_005, 90_004]  # consecutive lengths: every byte phase of the unaligned load path
    lens = np.concatenate([long_lens, rnd.integers(0, 400, size=3000)]).astype(np.int64)
    rnd.shuffle(lens)
    data, off = _ragged(rnd, lens)
    want = oracle.sha3_batch(data, off, d, threads=0)
    got = engine.sha3(data, off, d)
    assert np.array_equal(got, want), np.nonzero((got != want).any(axis=1))[0][:10]


def test_single_long_message_and_uniform_long_messages(engine, oracle):
    rnd = np.random.default_rng(41)
    for lens in ([1_000_000], [300_000] * 5, [72 * 4000] * 3, [0, 250_000]):
        data, off = _ragged(rnd, np.array(lens, dtype=np.int64))
        for d in (256, 512):
            assert np.array_equal(engine.sha3(data, off, d), oracle.sha3_batch(data, off, d, threads=0)), (lens, d)


@pytest.mark.parametrize("d", [224, 256, 512])
def test_chain_bound_kmac_and_cshake(engine, oracle, d):
    """KMACXOF / cSHAKE over long messages (key block + trailer boundary blocks, 172-byte rate quirk at D224)"""
    rnd = np.random.default_rng(42 + d)
    lens = np.array([150_000, 99_999, 64_000 + 7, 168 * 500, 136 * 700 - 3, 5, 0, 300, 172 * 400], dtype=np.int64)
    data, off = _ragged(rnd, lens)
    keys, koff = pack([bytes(rnd.integers(0, 256, size=k, dtype=np.uint8)) for k in (32, 0, 56, 200, 1, 64, 32, 32, 7)])
    for out_bits in (512, 8 * 300):
        got = engine.kmac_xof(keys, koff, data, off, out_bits, b"My Tagged Application", d)
        want = oracle.kmac_xof_batch(keys, koff, data, off, out_bits, b"My Tagged Application", d, threads=0)
        assert np.array_equal(got, want), (d, out_bits, np.nonzero((got != want).any(axis=1))[0])
    for n_str, s_str in ((b"", b"Email Signature"), (b"", b"")):  # the second is the double-padding quirk Q4
        got = engine.cshake(data, off, 512, n_str, s_str, d)
        want = oracle.cshake_batch(data, off, 512, n_str, s_str, d, threads=0)
        assert np.array_equal(got, want), (d, n_str, s_str)


def test_long_keystreams_through_the_pair_tier(engine):
    """sha3_encrypt of a few long messages: the keystream pass squeezes |m| bytes per item into the XOR (ragged by
    OUTPUT length), the tag pass absorbs the long plaintext; decrypt restores on a bad tag"""
    rnd = random.Random(43)
    msgs = [rnd.randbytes(n) for n in (120_000, 80_001, 60_002, 40_003, 17, 0, 136 * 300)]
    pws = [rnd.randbytes(16) for _ in msgs]
    nonces = [rnd.randbytes(512) for _ in msgs]
    pd, po = pack(pws)
    md, mo = pack(msgs)
    ct, tag = engine.sponge_encrypt(pd, po, b"".join(nonces), 512, md, mo, 512)
    for i, m in enumerate(msgs):
        ke_ka = R.kmac_xof(nonces[i] + pws[i], b"", 1024, b"S", 512)
        want_t = R.kmac_xof(ke_ka[64:], m, 512, b"SKA", 512) if len(m) < 2000 else None
        if want_t is not None:
            assert tag[i].tobytes() == want_t
    # independent check of the long items through the (already verified) KMAC entry point
    kd, ko = pack([R.kmac_xof(nonces[i] + pws[i], b"", 1024, b"S", 512)[:64] for i in range(len(msgs))])
    empty = np.zeros(len(msgs) + 1, np.uint64)
    ks = engine.kmac_xof(kd, ko, np.zeros(0, np.uint8), empty, 0, b"SKE", 512, out_off=mo)
    assert np.array_equal(ct, ks ^ md)
    out, ok = engine.sponge_decrypt(pd, po, b"".join(nonces), 512, ct, mo, tag, 512)
    assert ok.all() and np.array_equal(out, md)
    bad = ct.copy()
    bad[5] ^= 0x80
    out, ok = engine.sponge_decrypt(pd, po, b"".join(nonces), 512, bad, mo, tag, 512)
    assert ok.tolist() == [0] + [1] * (len(msgs) - 1)
    assert np.array_equal(out[:120_000], bad[:120_000]) and np.array_equal(out[120_000:], md[120_000:])

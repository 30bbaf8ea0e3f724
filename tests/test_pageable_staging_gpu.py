"""GPU: host entry points stage large PAGEABLE buffers through page-locked double buffers with a few copy threads
(csrc/ctx.cu: copy_in / copy_out / stage_drain); buffers from capy_host_alloc go straight to the DMA engine.  Both ways
must give the same bytes, for copies of many pieces through one staging slot, for outputs larger than a piece, and when a
call follows a call (the staging halves are reused).  Matches compute_sha3_hash (sha3/hashable.rs:19-21), kmac_xof
(sha3/shake_functions.rs:79-89), sha3_encrypt / sha3_decrypt (sha3/encryptable.rs:29-83)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_sha3_ragged_and_fixed_pageable_equals_pinned(engine, oracle):
    rnd = np.random.default_rng(41)
    # one ragged chunk of ~44 MB (long messages keep it in one chunk): six pieces through one staging slot
    lens = np.concatenate([rnd.integers(300_000, 900_000, size=60), rnd.integers(0, 4000, size=3000)])
    rnd.shuffle(lens)
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint64)
    data = rnd.integers(0, 256, size=int(off[-1]), dtype=np.uint8)
    pinned = engine.pinned(len(data))
    pinned[:] = data
    out_pinned = engine.pinned(len(lens) * 64).reshape(len(lens), 64)
    a = engine.sha3(data, off, 512)                        # pageable in, pageable out
    b = engine.sha3(pinned, off, 512, out=out_pinned)      # page-locked in and out
    c = engine.sha3(data, off, 512)                        # again: the staging halves are reused
    assert np.array_equal(a, b) and np.array_equal(a, c)
    pick = np.sort(rnd.choice(len(lens), size=40, replace=False))
    sm = [data[int(off[i]):int(off[i + 1])] for i in pick]
    so = np.concatenate([[0], np.cumsum([len(x) for x in sm])]).astype(np.uint64)
    assert np.array_equal(a[pick], oracle.sha3_batch(np.concatenate(sm), so, 512, threads=0))
    # fixed-length entry point: 2^19 x 64 B in 8 MiB chunks over three streams, 16 MiB of digests out
    n = 1 << 19
    d2 = rnd.integers(0, 256, size=n * 64, dtype=np.uint8)
    p2 = engine.pinned(n * 64)
    p2[:] = d2
    x = engine.sha3_fixed(d2, 64, 64, n, 256)
    y = engine.sha3_fixed(p2, 64, 64, n, 256, out=engine.pinned(n * 32).reshape(n, 32))
    assert np.array_equal(x, y)
    import hashlib

    for i in (0, 1, n // 2, n - 1):
        assert x[i].tobytes() == hashlib.sha3_256(d2[64 * i:64 * i + 64].tobytes()).digest()


def test_large_outputs_and_ae_pageable(engine, oracle):
    rnd = np.random.default_rng(42)
    # KMACXOF with 3 000-byte outputs: 24 MB out through the staging halves, several chunks
    n, mlen, ob = 8192, 1500, 3000
    data = rnd.integers(0, 256, size=n * mlen, dtype=np.uint8)
    off = np.arange(n + 1, dtype=np.uint64) * mlen
    keys = rnd.integers(0, 256, size=n * 32, dtype=np.uint8)
    koff = np.arange(n + 1, dtype=np.uint64) * 32
    a = engine.kmac_xof(keys, koff, data, off, 8 * ob, b"T", 512)
    pk, pd = engine.pinned(len(keys)), engine.pinned(len(data))
    pk[:], pd[:] = keys, data
    b = engine.kmac_xof(pk, koff, pd, off, 8 * ob, b"T", 512, out=engine.pinned(n * ob).reshape(n, ob))
    assert np.array_equal(a, b)
    assert np.array_equal(a[:8], oracle.kmac_xof_batch(keys[: 8 * 32], koff[:9], data[: 8 * mlen], off[:9], 8 * ob, b"T", 512))
    # sponge AE round trip from pageable memory: 12 MB of messages in, ciphertext out, and back
    pws = rnd.integers(0, 256, size=n * 16, dtype=np.uint8)
    po = np.arange(n + 1, dtype=np.uint64) * 16
    nonces = rnd.integers(0, 256, size=n * 512, dtype=np.uint8)
    ct, tag = engine.sponge_encrypt(pws, po, nonces, 512, data, off, 512)
    back, ok = engine.sponge_decrypt(pws, po, nonces, 512, ct, off, tag, 512)
    assert ok.all() and np.array_equal(back, data) and not np.array_equal(ct, data)

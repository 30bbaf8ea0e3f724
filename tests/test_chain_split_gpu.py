"""GPU: uniform cSHAKE / KMAC batches that the engine cuts into two dependent jobs (sponge_chain_kernel: the first job
absorbs half of every item's blocks and hands the state over, the second finishes) must give the bytes of the uncut
launch and of the oracle.  2^16 items on a 148-SM GPU is the shape that gets cut (BASELINE config 2).

Matches kmac_xof / cshake (sha3/shake_functions.rs:49-89) over equal-length messages."""
import ctypes as C
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

N = 1 << 16


def _would_cut(engine, absorb, squeeze_extra=0):
    import torch

    sms = torch.cuda.get_device_properties(0).multi_processor_count
    cut = C.c_uint64(0)
    return engine.lib.capy_chain_cut(sms, N, absorb, squeeze_extra, C.byref(cut)) == 1


def _both(fn):
    """fn() with the cut enabled and disabled (CAPY_NO_CHAIN_SPLIT is read at every launch)"""
    os.environ.pop("CAPY_NO_CHAIN_SPLIT", None)
    a = fn()
    os.environ["CAPY_NO_CHAIN_SPLIT"] = "1"
    try:
        b = fn()
    finally:
        os.environ.pop("CAPY_NO_CHAIN_SPLIT", None)
    return a, b


@pytest.mark.parametrize("d,mlen", [(512, 2040), (512, 2041), (512, 2176), (512, 2177), (256, 2520), (256, 2600), (384, 2500)])
def test_kmac_fixed_cut_equals_uncut_and_oracle(engine, oracle, d, mlen):
    rate = (1600 - d) // 8
    assert _would_cut(engine, 1 + mlen // rate + 1), "the shape under test is supposed to be cut on this GPU"
    g = torch.Generator(device="cuda")
    g.manual_seed(d + mlen)
    data = torch.randint(0, 256, (N * mlen,), dtype=torch.uint8, device="cuda", generator=g)
    keys = torch.randint(0, 256, (N * 32,), dtype=torch.uint8, device="cuda", generator=g)

    def run():
        out = torch.zeros(N * 64, dtype=torch.uint8, device="cuda")
        engine.kmac_xof_fixed_dev(keys, 32, 32, data, mlen, mlen, N, 512, b"My Tagged Application", d, out)
        torch.cuda.synchronize()
        return out

    cut, plain = _both(run)
    assert torch.equal(cut, plain)
    pick = np.sort(np.random.default_rng(mlen).choice(N, size=48, replace=False))
    idx = torch.from_numpy(pick).cuda()
    s_data = data.view(N, mlen)[idx].cpu().numpy().reshape(-1)
    s_keys = keys.view(N, 32)[idx].cpu().numpy().reshape(-1)
    want = oracle.kmac_xof_batch(s_keys, np.arange(49, dtype=np.uint64) * 32, s_data, np.arange(49, dtype=np.uint64) * mlen, 512,
                                 b"My Tagged Application", d)
    assert np.array_equal(cut.view(N, 64)[idx].cpu().numpy(), want)


@pytest.mark.parametrize("out_bytes", [64, 2000])
def test_cshake_and_kmac_with_offsets_and_long_outputs(engine, oracle, out_bytes):
    """offsets arrays on the device (the plan finds one block count), 2 000-byte outputs: the cut lies at the end of the absorb"""
    mlen = 2100
    g = torch.Generator(device="cuda")
    g.manual_seed(out_bytes)
    data = torch.randint(0, 256, (N * mlen,), dtype=torch.uint8, device="cuda", generator=g)
    keys = torch.randint(0, 256, (N * 40,), dtype=torch.uint8, device="cuda", generator=g)
    off = torch.arange(N + 1, dtype=torch.int64, device="cuda") * mlen
    koff = torch.arange(N + 1, dtype=torch.int64, device="cuda") * 40

    def cshake():
        out = torch.zeros(N * out_bytes, dtype=torch.uint8, device="cuda")
        engine.cshake_dev(data, off, 8 * out_bytes, b"", b"Email Signature", 512, out)
        torch.cuda.synchronize()
        return out

    def kmac():
        out = torch.zeros(N * out_bytes, dtype=torch.uint8, device="cuda")
        engine.kmac_xof_dev(keys, koff, data, off, 8 * out_bytes, b"tag", 512, out)
        torch.cuda.synchronize()
        return out

    pick = np.sort(np.random.default_rng(out_bytes).choice(N, size=24, replace=False))
    idx = torch.from_numpy(pick).cuda()
    s_data = data.view(N, mlen)[idx].cpu().numpy().reshape(-1)
    s_off = np.arange(25, dtype=np.uint64) * mlen
    c_cut, c_plain = _both(cshake)
    assert torch.equal(c_cut, c_plain)
    assert np.array_equal(c_cut.view(N, out_bytes)[idx].cpu().numpy(),
                          oracle.cshake_batch(s_data, s_off, 8 * out_bytes, b"", b"Email Signature", 512))
    k_cut, k_plain = _both(kmac)
    assert torch.equal(k_cut, k_plain)
    s_keys = keys.view(N, 40)[idx].cpu().numpy().reshape(-1)
    assert np.array_equal(k_cut.view(N, out_bytes)[idx].cpu().numpy(),
                          oracle.kmac_xof_batch(s_keys, np.arange(25, dtype=np.uint64) * 40, s_data, s_off, 8 * out_bytes, b"tag", 512))


def test_host_entry_point_with_ragged_keys(engine, oracle):
    """capy_kmac_xof_batch: equal-length messages (found from the host offsets), keys of different lengths -- some items
    have one key block, some two, so the cut is not at the same place of every stream; still the same bytes"""
    mlen, n = 1800, N
    rnd = np.random.default_rng(9)
    data = rnd.integers(0, 256, size=n * mlen, dtype=np.uint8)
    off = np.arange(n + 1, dtype=np.uint64) * mlen
    klens = rnd.choice([0, 16, 32, 131, 140, 267], size=n)
    koff = np.concatenate([[0], np.cumsum(klens)]).astype(np.uint64)
    keys = rnd.integers(0, 256, size=int(koff[-1]), dtype=np.uint8)
    cut, plain = _both(lambda: engine.kmac_xof(keys, koff, data, off, 512, b"", 512))
    assert np.array_equal(cut, plain)
    pick = np.sort(rnd.choice(n, size=64, replace=False))
    s_keys = np.concatenate([keys[int(koff[i]):int(koff[i + 1])] for i in pick])
    s_koff = np.concatenate([[0], np.cumsum(klens[pick])]).astype(np.uint64)
    s_data = data.reshape(n, mlen)[pick].reshape(-1)
    assert np.array_equal(cut[pick], oracle.kmac_xof_batch(s_keys, s_koff, s_data, np.arange(65, dtype=np.uint64) * mlen, 512, b"", 512))


def test_open_as_dependent_jobs(engine):
    """sha3_decrypt (sha3/encryptable.rs:58-83) over 2^16 ragged messages: the keystream pass and the tag pass over the
    recovered plaintext run as dependent jobs of one launch; same plaintexts and verdicts as two launches, tampered
    items are caught and get their ciphertext back, and a few items are checked against the Python restatement."""
    from oracle import ref_sha3 as R

    n = N
    rnd = np.random.default_rng(31)
    lens = rnd.integers(0, 700, size=n)
    lens[:4] = [0, 136, 137, 699]
    mo = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    msg = rnd.integers(0, 256, size=int(mo[-1]), dtype=np.uint8)
    pw = rnd.integers(0, 256, size=n * 16, dtype=np.uint8)
    nonces = rnd.integers(0, 256, size=n * 512, dtype=np.uint8)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    t_msg, t_mo, t_pw, t_po, t_n = t(msg), t(mo), t(pw), t(np.arange(n + 1, dtype=np.int64) * 16), t(nonces)
    ct = torch.zeros_like(t_msg)
    tag = torch.zeros(n * 64, dtype=torch.uint8, device="cuda")
    engine.sponge_encrypt_dev(t_pw, t_po, n * 16, t_n, 512, t_msg, t_mo, 512, ct, tag)
    torch.cuda.synchronize()
    for i in (0, 1, 2, 3, n - 1):
        c_ref, t_ref = R.sha3_encrypt(msg[int(mo[i]):int(mo[i + 1])].tobytes(), pw[16 * i:16 * i + 16].tobytes(), 512,
                                      nonces[512 * i:512 * i + 512].tobytes())
        assert ct[int(mo[i]):int(mo[i + 1])].cpu().numpy().tobytes() == c_ref and tag[64 * i:64 * i + 64].cpu().numpy().tobytes() == t_ref
    bad = [7, n // 2, n - 2]
    tag2 = tag.clone()
    for i in bad:
        tag2[64 * i + 5] ^= 1

    def open_(tg):
        def run():
            out = torch.zeros_like(t_msg)
            ok = torch.zeros(n, dtype=torch.uint8, device="cuda")
            engine.sponge_decrypt_dev(t_pw, t_po, n * 16, t_n, 512, ct, t_mo, tg, 512, out, ok)
            torch.cuda.synchronize()
            return out, ok
        return run

    (o1, k1), (o2, k2) = _both(open_(tag))
    assert torch.equal(o1, o2) and torch.equal(k1, k2) and bool(k1.all().item()) and torch.equal(o1, t_msg)
    (o1, k1), (o2, k2) = _both(open_(tag2))
    assert torch.equal(o1, o2) and torch.equal(k1, k2)
    ok = k1.cpu().numpy()
    assert int(ok.sum()) == n - len(bad) and not ok[bad].any()
    o = o1.cpu().numpy()
    c = ct.cpu().numpy()
    for i in bad:  # restore-on-failure: the ciphertext comes back (encryptable.rs:77-82)
        assert np.array_equal(o[int(mo[i]):int(mo[i + 1])], c[int(mo[i]):int(mo[i + 1])])


@pytest.mark.parametrize("d,mlen", [(256, 2000), (512, 1000), (224, 2015), (384, 1400)])
def test_sha3_equal_long_messages(engine, oracle, d, mlen):
    """compute_sha3_hash (sha3/hashable.rs:19-21) over 2^16 equal messages of a dozen blocks or more: fixed-length entry
    point and offsets entry point, cut and uncut, against the oracle (the quirk lengths of Q1 / Q2 included via mlen)"""
    rate = {224: 144, 256: 136, 384: 104, 512: 72}[d]
    assert _would_cut(engine, mlen // rate + 1)
    g = torch.Generator(device="cuda")
    g.manual_seed(3 * d + mlen)
    data = torch.randint(0, 256, (N * mlen,), dtype=torch.uint8, device="cuda", generator=g)
    off = torch.arange(N + 1, dtype=torch.int64, device="cuda") * mlen

    def fixed():
        out = torch.zeros(N * (d // 8), dtype=torch.uint8, device="cuda")
        engine.sha3_fixed_dev(data, mlen, mlen, N, d, out)
        torch.cuda.synchronize()
        return out

    def ragged():
        out = torch.zeros(N * (d // 8), dtype=torch.uint8, device="cuda")
        engine.sha3_dev(data, off, d, out)
        torch.cuda.synchronize()
        return out

    f_cut, f_plain = _both(fixed)
    r_cut, r_plain = _both(ragged)
    assert torch.equal(f_cut, f_plain) and torch.equal(r_cut, r_plain) and torch.equal(f_cut, r_cut)
    pick = np.sort(np.random.default_rng(d).choice(N, size=64, replace=False))
    idx = torch.from_numpy(pick).cuda()
    want = oracle.sha3_batch(data.view(N, mlen)[idx].cpu().numpy().reshape(-1), np.arange(65, dtype=np.uint64) * mlen, d, threads=0)
    assert np.array_equal(f_cut.view(N, d // 8)[idx].cpu().numpy(), want)


@pytest.mark.parametrize("bits,mlen,out_bytes", [(256, 3000, 64), (128, 3000, 500), (256, 2000, 2000)])
def test_fips_shake_cut_and_ragged(engine, bits, mlen, out_bytes):
    """FIPS 202 SHAKE128 / SHAKE256 (no reference counterpart; checked against hashlib): a uniform batch that is cut, and a
    ragged one that is ordered longest first"""
    import hashlib

    g = torch.Generator(device="cuda")
    g.manual_seed(bits + mlen)
    data = torch.randint(0, 256, (N * mlen,), dtype=torch.uint8, device="cuda", generator=g)
    off = torch.arange(N + 1, dtype=torch.int64, device="cuda") * mlen

    def run():
        out = torch.zeros(N * out_bytes, dtype=torch.uint8, device="cuda")
        engine.fips_shake_dev(data, off, bits, out_bytes, out)
        torch.cuda.synchronize()
        return out

    cut, plain = _both(run)
    assert torch.equal(cut, plain)
    h = hashlib.shake_256 if bits == 256 else hashlib.shake_128
    for i in (0, 1, N // 3, N - 1):
        m = data[i * mlen:(i + 1) * mlen].cpu().numpy().tobytes()
        assert cut[i * out_bytes:(i + 1) * out_bytes].cpu().numpy().tobytes() == h(m).digest(out_bytes)
    # ragged: lengths 0..5000 (sorted launch), same check
    rnd = np.random.default_rng(bits)
    lens = rnd.integers(0, 5000, size=3000)
    o2 = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    d2 = data[: int(o2[-1])]
    out2 = torch.zeros(3000 * out_bytes, dtype=torch.uint8, device="cuda")
    engine.fips_shake_dev(d2, torch.from_numpy(o2).cuda(), bits, out_bytes, out2)
    torch.cuda.synchronize()
    for i in (0, 7, 1500, 2999):
        m = d2[int(o2[i]):int(o2[i + 1])].cpu().numpy().tobytes()
        assert out2[i * out_bytes:(i + 1) * out_bytes].cpu().numpy().tobytes() == h(m).digest(out_bytes)

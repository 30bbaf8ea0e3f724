"""CPU: when the engine cuts the chains of a uniform cSHAKE / KMAC batch in two dependent jobs (capy_chain_cut, the
planner behind csrc/sponge.cuh: sponge_chain_kernel).  Matches kmac_xof / cshake over equal-length messages
(sha3/shake_functions.rs:49-89): same bytes, another launch shape."""
import ctypes as C

from capycrypt_b200 import _binding as B


def _cut(n, absorb, squeeze_extra=0, sms=148):
    lib = B.load()
    cut = C.c_uint64(0)
    rc = lib.capy_chain_cut(sms, n, absorb, squeeze_extra, C.byref(cut))
    assert rc in (0, 1)
    return rc, cut.value


def test_cfg2_is_cut_in_half():
    # 2^16 items on 148 SMs: 3.46 warps per scheduler -> 6.92 after the cut; 32 blocks after the cached prefix
    assert _cut(1 << 16, 32) == (1, 16)
    # 4 KB in, 4 KB out (cSHAKE: 31 absorbed blocks, 30 permutations between squeeze blocks): cut at the end of the absorb
    assert _cut(1 << 16, 31, 30) == (1, 31)
    # keystream shape (2 absorbed blocks, 30 squeeze permutations): the absorb phase cannot carry a half
    assert _cut(1 << 16, 2, 30)[0] == 0


def test_batches_that_are_left_alone():
    assert _cut(1 << 17, 32)[0] == 0  # 6.92 warps per scheduler already
    assert _cut(1 << 20, 32)[0] == 0  # large batch: the partial wave does not matter
    assert _cut(50_000, 32)[0] == 0   # fewer job-0 blocks than the GPU holds at once: job 1 would wait
    assert _cut(1 << 16, 8)[0] == 0   # short chains
    assert _cut(1 << 16, 32, sms=0)[0] == 0
    # 3 x 592 x 32 items fill every scheduler with exactly three warps: nothing to gain
    assert _cut(3 * 592 * 32, 32)[0] == 0

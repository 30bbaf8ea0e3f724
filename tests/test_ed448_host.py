"""CPU suite: the engine's own __host__ __device__ Ed448 arithmetic (capycrypt_b200/csrc/{fp448,sc448,
ed448}.cuh -- the code the CUDA kernels run) compiled for the CPU by tests/host/ed448_host_check.cpp and
checked against the oracle.  This is how the device logic is validated in the GPU-less build container;
the GPU tests (tests/test_ed448_gpu.py) are the parity tests proper."""
import ctypes as C
import os
import random
import subprocess

import pytest

from oracle import ref_ed448 as E

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "host", "ed448_host_check.cpp")
LIB = os.path.join(HERE, "host", "_build", "libed448_host.so")
P, R = E.P, E.R


@pytest.fixture(scope="module")
def host():
    deps = [SRC] + [os.path.join(HERE, "..", "capycrypt_b200", "csrc", f) for f in ("fp448.cuh", "sc448.cuh", "ed448.cuh")]
    if not os.path.exists(LIB) or os.path.getmtime(LIB) < max(os.path.getmtime(d) for d in deps):
        os.makedirs(os.path.dirname(LIB), exist_ok=True)
        subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-Wno-unknown-pragmas", "-o", LIB, SRC], check=True)
    return C.CDLL(LIB)


def _fe(host, op, a, b=0):
    o = (C.c_uint8 * 56)()
    host.host_fe_op(op, a.to_bytes(56, "little"), b.to_bytes(56, "little"), o)
    return int.from_bytes(bytes(o), "little")


def _sc(host, op, a, b=0):
    o = (C.c_uint8 * 56)()
    host.host_sc_op(op, a.to_bytes(56, "big"), b.to_bytes(56, "big"), o)
    return int.from_bytes(bytes(o), "big")


EDGE_FE = [0, 1, 2, P - 1, P - 2, 2**224, 2**224 - 1, 2**448 - 1, P, P + 1, (1 << 448) - (1 << 224), 2**28 - 1,
           int("fffffff" * 16, 16), int("0000000fffffff" * 8, 16)]


def test_field_ops(host):
    rnd = random.Random(1)
    vals = EDGE_FE + [rnd.randrange(2**448) for _ in range(300)]
    for a in vals:
        b = rnd.choice(vals)
        assert _fe(host, 0, a, b) == a * b % P
        assert _fe(host, 1, a) == a * a % P
        assert _fe(host, 3, a, b) == (a + b) % P
        assert _fe(host, 4, a, b) == (a - b) % P
        assert _fe(host, 5, a, b) == (a + b) * (a - b) % P  # loose (unreduced) operands into the multiplier


def test_field_ops_worst_case_limbs(host):
    """All-ones 28-bit limbs (every limb at its maximum) through the multiplier at the loosest operand bounds the
    point formulas use (2 x 3): the 64-bit column sums and the in-order carry chains must not overflow."""
    ones = int("fffffff" * 16, 16)
    hi = (1 << 448) - 1
    for a in (ones, hi, P - 1, ones - 1):
        for b in (ones, hi, 0, 1, P - 1):
            assert _fe(host, 0, a, b) == a * b % P
            assert _fe(host, 1, a) == a * a % P
            assert _fe(host, 5, a, b) == (a + b) * (a - b) % P


def test_field_inverse(host):
    rnd = random.Random(2)
    for a in EDGE_FE + [rnd.randrange(P) for _ in range(20)]:
        assert _fe(host, 2, a) == pow(a, P - 2, P)


def test_scalar_ops(host):
    rnd = random.Random(3)
    edge = [0, 1, R - 1, R, R + 1, 2**446, 2**446 - 1, 2**448 - 1, 4 * R - 1, (R + 1) // 4]
    vals = edge + [rnd.randrange(2**448) for _ in range(300)]
    for a in vals:
        b = rnd.choice(vals)
        assert _sc(host, 0, a, b) == a * b % R
        assert _sc(host, 1, a) == 4 * a % R
        assert _sc(host, 2, a, b) == (a - b) % R
        assert _sc(host, 3, a) == a % R


def _phi(pt):
    """4-isogeny E -> E' (twisted, a = -1, d' = d - 1): (2xy / (y^2 - x^2), (y^2 + x^2) / (2 - y^2 - x^2))."""
    x, y = pt
    return (2 * x * y * E.inv((y * y - x * x) % P) % P, (y * y + x * x) * E.inv((2 - y * y - x * x) % P) % P)


def test_comb_table_entries(host):
    """Entry (i, j) of the fixed-base table is (j+1) * 32^i * phi(G) as (y - x, y + x, 2 d' x y), d' = -39082."""
    o = (C.c_uint8 * 168)()
    d_tw = (E.D - 1) % P
    for i, j in [(0, 0), (0, 15), (1, 0), (5, 3), (60, 4), (89, 0), (89, 15)]:
        host.host_table_entry(i, j, o)
        b = bytes(o)
        x, y = _phi(E.scalar_mult((j + 1) * 32**i, E.GENERATOR))
        assert (-x * x + y * y - 1 - d_tw * x * x * y * y) % P == 0
        assert int.from_bytes(b[:56], "little") == (y - x) % P
        assert int.from_bytes(b[56:112], "little") == (y + x) % P
        assert int.from_bytes(b[112:], "little") == 2 * d_tw * x * y % P


SCALARS = [0, 1, 2, 7, 8, 9, 15, 16, R - 1, R, R + 1, 2**446 - 1, 2**448 - 1, int("8" * 112, 16), int("7" * 112, 16),
           int("f" * 112, 16), int("9" * 112, 16)]


def test_fixed_base(host):
    rnd = random.Random(4)
    out = (C.c_uint8 * 112)()
    for k in SCALARS + [rnd.randrange(2**448) for _ in range(24)]:
        for ct in (0, 1):
            host.host_fixed_base(k.to_bytes(56, "big"), ct, out)
            assert bytes(out) == E.point_to_bytes(E.scalar_mult(k % R, E.GENERATOR)), (k, ct)


def test_var_base_including_small_order_and_unreduced(host):
    rnd = random.Random(5)
    out = (C.c_uint8 * 112)()
    pts = [E.IDENTITY, (0, P - 1), (1, 0), (P - 1, 0), E.GENERATOR, E.point_add(E.GENERATOR, (1, 0))]
    pts += [E.scalar_mult(rnd.randrange(R), E.GENERATOR) for _ in range(3)]
    for pt in pts:
        for k in rnd.sample(SCALARS, 5) + [2**448 - 1, 0, 1, R, rnd.randrange(2**448)]:
            for ct in (0, 1):
                assert host.host_var_base(k.to_bytes(56, "big"), E.point_to_bytes(pt), ct, out) == 0
                assert bytes(out) == E.point_to_bytes(E.scalar_mult(k, pt)), (k, pt, ct)
    assert host.host_var_base((5).to_bytes(56, "big"), E.point_to_bytes((5, 7)), 1, out) == -4


def test_var_base_with_addend(host):
    """The verify shape U = [z]G + [h]V (ecc/signable.rs:77): the addend is the last iteration of the ladder loop."""
    rnd = random.Random(6)
    out = (C.c_uint8 * 112)()
    for _ in range(6):
        pt = E.scalar_mult(rnd.randrange(R), E.GENERATOR)
        add = rnd.choice([E.IDENTITY, pt, E.scalar_mult(rnd.randrange(R), E.GENERATOR), (0, P - 1)])
        k = rnd.choice([0, 1, 2**448 - 1, rnd.randrange(2**448)])
        for ct in (0, 1):
            assert host.host_var_base2(k.to_bytes(56, "big"), E.point_to_bytes(pt), ct, E.point_to_bytes(add), out) == 0
            assert bytes(out) == E.point_to_bytes(E.point_add(E.scalar_mult(k, pt), add)), (k, pt, add, ct)

"""GPU: the scalar / field helpers give the same bytes on the device as the same source compiled for the
CPU (guards against codegen differences; an earlier sc_mul_mod formulation was miscompiled for sm_100a)."""
import json
import os
import subprocess

import pytest

pytestmark = pytest.mark.gpu
EXE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "host", "_build", "device_selfcheck")


def test_device_matches_host_for_field_and_scalar_helpers():
    assert os.path.exists(EXE), "run __graft_entry__.build() first"
    r = subprocess.run([EXE], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    res = json.loads(r.stdout.strip().splitlines()[-1])
    assert all(v == 0 for v in res["mismatch"].values()), res

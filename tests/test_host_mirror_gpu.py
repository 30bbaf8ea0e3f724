"""GPU: the C++ host mirror of the reference's operator surface (capycrypt_b200/host/capycrypt_gpu.hpp) reproduces
the reference's own KATs and round-trip tests through the C ABI."""
import os
import subprocess

import pytest

pytestmark = pytest.mark.gpu
EXE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "host", "_build", "host_mirror_check")


def test_cpp_host_mirror():
    assert os.path.exists(EXE), "run __graft_entry__.build() first"
    env = dict(os.environ)
    libdir = os.path.join(os.path.dirname(EXE), "..", "..", "..", "capycrypt_b200", "_lib")
    env["LD_LIBRARY_PATH"] = os.path.abspath(libdir) + ":" + env.get("LD_LIBRARY_PATH", "")
    r = subprocess.run([EXE], capture_output=True, text=True, timeout=300, env=env)
    assert r.returncode == 0 and "host mirror ok" in r.stdout, r.stdout + r.stderr

"""CPU: host-side batch plumbing of the C ABI layer (capycrypt_b200/csrc/hostbatch.h) -- how batches are split over
devices and into pipeline chunks.  Built with g++ against the CUDA headers; no device is touched."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))


def test_hostbatch_check(tmp_path):
    exe = str(tmp_path / "hostbatch_check")
    cuda_inc = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "include")
    subprocess.run(["g++", "-O1", "-std=c++17", "-I" + cuda_inc, "-o", exe, os.path.join(HERE, "host", "hostbatch_check.cpp")],
                   check=True)
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0 and "hostbatch ok" in r.stdout, r.stdout + r.stderr

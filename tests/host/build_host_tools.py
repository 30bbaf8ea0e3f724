"""Builds the test scaffolding under tests/host/_build (never part of the product library):
  libed448_host.so   the engine's __host__ __device__ Ed448 code compiled for the CPU (g++)
  device_selfcheck   the same helpers run on GPU and CPU side by side (nvcc, sm_100a)"""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_build")


def build(force: bool = False) -> None:
    os.makedirs(OUT, exist_ok=True)
    csrc = os.path.join(HERE, "..", "..", "capycrypt_b200", "csrc")
    deps = [os.path.join(csrc, f) for f in ("fp448.cuh", "sc448.cuh", "ed448.cuh")]

    def stale(target, srcs):
        return force or not os.path.exists(target) or os.path.getmtime(target) < max(os.path.getmtime(s) for s in srcs)

    lib = os.path.join(OUT, "libed448_host.so")
    src = os.path.join(HERE, "ed448_host_check.cpp")
    if stale(lib, deps + [src]):
        subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-Wno-unknown-pragmas", "-o", lib, src], check=True)
    exe = os.path.join(OUT, "device_selfcheck")
    src = os.path.join(HERE, "device_selfcheck.cu")
    if stale(exe, deps + [src]):
        subprocess.run(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "--expt-relaxed-constexpr",
                        "-o", exe, src], check=True)


    # C++ host mirror check (links the product library; runs only on a GPU box)
    root = os.path.join(HERE, "..", "..")
    libdir = os.path.join(root, "capycrypt_b200", "_lib")
    exe = os.path.join(OUT, "host_mirror_check")
    src = os.path.join(HERE, "host_mirror_check.cpp")
    hdrs = [os.path.join(root, "capycrypt_b200", "host", "capycrypt_gpu.hpp"), os.path.join(root, "include", "capy_gpu.h")]
    if os.path.exists(os.path.join(libdir, "libcapycrypt_gpu.so")) and stale(exe, hdrs + [src]):
        subprocess.run(["g++", "-O2", "-std=c++17", "-o", exe, src, "-L" + libdir, "-lcapycrypt_gpu",
                        "-Wl,-rpath," + os.path.abspath(libdir)], check=True)


if __name__ == "__main__":
    build()

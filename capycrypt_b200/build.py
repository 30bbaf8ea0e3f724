"""Builds libcapycrypt_gpu.so in-tree with nvcc for sm_100a (no GPU needed: cross-compiles).

    python -m capycrypt_b200.build [--force] [-D NAME=VALUE ...]

The shared library lands in capycrypt_b200/_lib/ (git-ignored; it travels to the GPU box with
the gpurun snapshot).
"""
from __future__ import annotations

import concurrent.futures as cf
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "_lib")
LIB = os.path.join(LIBDIR, "libcapycrypt_gpu.so")
SOURCES = ["ctx.cu", "sha3_api.cu", "ed448_api.cu", "ed448_fixed.cu", "ed448_var.cu", "ae_api.cu"]
# per-source defines (none at present: both curve kernels are fastest with every field multiplication inlined --
# csrc/ed448_var.cu and csrc/ed448_fixed.cu explain their code shape)
PER_SOURCE_DEFINES = {}
NVCC = os.environ.get("NVCC", "nvcc")
BASE_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC,-O3,-Wall,-Wno-unknown-pragmas", "--expt-relaxed-constexpr",
]


def _deps():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC)) + [os.path.join(HERE, "..", "include", "capy_gpu.h")]


def _stamp(defines):
    h = hashlib.sha256()
    for p in _deps():
        h.update(p.encode())
        h.update(open(p, "rb").read())
    h.update(" ".join(BASE_FLAGS + defines).encode())
    h.update(repr(sorted(PER_SOURCE_DEFINES.items())).encode())
    return h.hexdigest()


def build(force: bool = False, defines=(), verbose: bool = False, lib: str = LIB) -> str:
    defines = [f"-D{d}" for d in defines]
    os.makedirs(LIBDIR, exist_ok=True)
    stamp_file = lib + ".stamp"
    stamp = _stamp(defines)
    if not force and os.path.exists(lib) and os.path.exists(stamp_file) and open(stamp_file).read() == stamp:
        return lib
    objdir = os.path.join(LIBDIR, "obj_" + hashlib.sha256(lib.encode()).hexdigest()[:8])
    os.makedirs(objdir, exist_ok=True)

    def compile_one(src):
        obj = os.path.join(objdir, src.replace(".cu", ".o"))
        cmd = [NVCC, *BASE_FLAGS, *PER_SOURCE_DEFINES.get(src, []), *defines, "-Xptxas", "-v", "-c",
               os.path.join(CSRC, src), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
        with open(obj + ".ptxas.log", "w") as f:
            f.write(r.stderr)
        return obj

    with cf.ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    cmd = [NVCC, "-shared", "-o", lib, *objs, "-gencode", "arch=compute_100a,code=sm_100a", "-cudart", "static"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    # standalone measurement tools: integer-pipe peaks (roofline denominators), Keccak unroll/block-size sweep,
    # chain speed of the split-lane / 25-thread permutations (profiles/r01*_keccak_*.jsonl)
    for tool in ("peaks", "keccak_sweep", "keccak_pair_probe", "fe_mul_probe"):
        r = subprocess.run([NVCC, *BASE_FLAGS, os.path.join(CSRC, tool + ".cu"), "-o", os.path.join(LIBDIR, tool),
                            "-cudart", "static"], capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"{tool} build failed:\n{r.stdout}\n{r.stderr}")
    with open(stamp_file, "w") as f:
        f.write(stamp)
    if verbose:
        for o in objs:
            sys.stdout.write(open(o + ".ptxas.log").read())
    return lib


if __name__ == "__main__":
    args = sys.argv[1:]
    defs = [a[2:] for a in args if a.startswith("-D")]
    out = build(force="--force" in args, defines=defs, verbose="-v" in args)
    print(out)

"""capycrypt_b200 -- B200 (sm_100a) batch engine for capyCRYPT's SHA3/KMAC and Ed448 hot paths.

The CUDA shared library (capycrypt_b200/_lib/libcapycrypt_gpu.so, C ABI in
include/capy_gpu.h) is the product; this package is the thin Python front end over it.
Nothing here computes a digest or a curve point on the CPU: without the library, import of
the engine fails.
"""
__all__ = ["Engine", "pack"]


def __getattr__(name):
    if name in ("Engine", "pack"):
        from . import engine

        return getattr(engine, name)
    raise AttributeError(name)

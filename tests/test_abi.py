"""CPU suite: the C-ABI library exists, loads, and exports every symbol include/capy_gpu.h declares
(no compute calls -- there is no GPU here)."""
import ctypes
import os

from capycrypt_b200 import _binding as B


def test_header_and_binding_agree():
    declared = set(B.declared_symbols())
    bound = set(B.SIGNATURES)
    assert declared == bound, (declared - bound, bound - declared)


def test_library_exports_every_declared_symbol():
    assert os.path.exists(B.LIB_PATH), "build with python -m capycrypt_b200.build (or __graft_entry__.build())"
    lib = ctypes.CDLL(B.LIB_PATH)
    for name in B.declared_symbols():
        assert hasattr(lib, name), name


def test_strerror_and_version_without_gpu():
    lib = B.load()
    assert lib.capy_version() >= 100
    assert b"security parameter" in lib.capy_strerror(B.ERR_BAD_SECPARAM)
    assert lib.capy_strerror(0) == b"ok"


def test_no_cpu_fallback_when_library_missing(tmp_path):
    import pytest

    with pytest.raises(ImportError):
        B.load(str(tmp_path / "libmissing.so"))


def test_product_never_imports_oracle():
    """The product path must not route through the oracle (or any CPU fallback)."""
    root = os.path.dirname(os.path.abspath(B.__file__))
    for dirpath, _, files in os.walk(root):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                src = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "import oracle" not in src and "from oracle" not in src and "ref_cpu" not in src, f
                assert "libcapy_oracle" not in src, f

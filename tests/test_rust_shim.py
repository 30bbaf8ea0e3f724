"""CPU: the Rust shim (rust/gpu.rs) has never been through rustc in this image, so its `extern "C"` block is checked here
against include/capy_gpu.h mechanically: same function names, same arity, same parameter names, and for every parameter
and return value the Rust type the C type maps to (pointer constness, integer width).  tools/gen_rust_extern.py holds the
two parsers and generates the block."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import gen_rust_extern as G  # noqa: E402

from capycrypt_b200 import _binding as B  # noqa: E402


def test_rust_extern_block_binds_every_export_with_the_header_types():
    c, r = G.parse_header(), G.parse_rust()
    assert set(c) == set(B.declared_symbols()), "the header parser missed a declaration"
    assert set(r) == set(c), (set(c) - set(r), set(r) - set(c))
    for name in c:
        c_ret, c_params = c[name]
        r_ret, r_params = r[name]
        assert r_ret == c_ret, (name, r_ret, c_ret)
        assert len(r_params) == len(c_params), name
        for (cn, ct), (rn, rt) in zip(c_params, r_params):
            assert (cn, ct) == (rn, rt), (name, cn, ct, rn, rt)


def test_type_mapping():
    assert G.c_type_to_rust("const uint8_t*") == "*const u8"
    assert G.c_type_to_rust("capy_ctx**") == "*mut *mut CapyCtx"
    assert G.c_type_to_rust("const capy_ctx*") == "*const CapyCtx"
    assert G.c_type_to_rust("uint64_t") == "u64" and G.c_type_to_rust("size_t") == "usize" and G.c_type_to_rust("int") == "c_int"


def test_shim_uses_the_c_abi_strides_it_checks():
    """The two ADVICE r1 findings stay fixed: sha3_decrypt passes the group's own nonce length (not a literal 512) and
    verify keeps malformed signatures out of the fixed-stride batch."""
    src = open(G.RUST).read()
    body = src[src.index("pub fn sha3_decrypt"):src.index("pub fn key_encrypt")]
    assert "nonce_len as u64" in body and "z.as_ptr(), 512" not in body
    body = src[src.index("pub fn verify"):src.index("pub fn sha3_encrypt")]
    assert "h.len() == 56" in body and "debug_assert!" in body

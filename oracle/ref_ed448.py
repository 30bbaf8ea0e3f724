"""CPU oracle (TEST INFRASTRUCTURE ONLY) -- Python big-int restatement of capyCRYPT's Ed448 path.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this
module.  The product path (capycrypt_b200/) never does.

What the reference does on this path is protocol glue (src/ecc/keypair.rs:41-51,
src/ecc/signable.rs:40-86, src/ecc/encryptable.rs:34-94); the curve arithmetic itself
lives in the third-party crate `tiny_ed448_goldilocks` pinned at 0.1.8
(Cargo.toml:22, Cargo.lock:857-868; transitively fiat-crypto 0.2.9, crypto-bigint 0.5.5),
which is NOT vendored under /root/reference and cannot be fetched (no network, no cargo).
Its published algorithm is restated here from the mathematics: the untwisted Edwards
curve  x^2 + y^2 = 1 + d x^2 y^2  over GF(2^448 - 2^224 - 1), d = -39081, prime subgroup
order r, RFC 8032 base point; complete unified addition.

PARITY STATUS: **parity unpinned** against the Rust crate -- the reference's tests hold
no known-answer vector for public keys, signatures or points (only random round trips,
tests/integration_tests.rs:20-130).  What IS pinned, independently:
  * the Keccak/KMAC side by the reference's NIST KATs (oracle/ref_sha3.py);
  * field + group arithmetic + base point by OpenSSL: for random seeds the RFC 8032
    Ed448 public key (57-byte encoding of [s]B, both coordinates) and RFC 7748 X448
    (u = y^2/x^2 of [k]P) agree with this module (tests/test_oracle_ed448.py);
  * group laws: r*G = O, (a+b)G = aG + bG, a(bG) = b(aG);
  * the reference's own self-consistency tests restated: verify(sign(m)) accepts,
    decrypt(encrypt(m)) round-trips, wrong key rejects.
Unverifiable crate details are single-point switches: GENERATOR,
`Scalar * Scalar` reducing mod r, canonical z in [0, r).
"""
from __future__ import annotations

from . import ref_sha3

P = 2**448 - 2**224 - 1
D = (-39081) % P
R = 2**446 - 13818066809895115352007386748515426880336692474882178609894547503885

# RFC 8032 section 5.2 base point (best-evidence generator).
GX = 0x4F1970C66BED0DED221D15A622BF36DA9E146570470F1767EA6DE324A3D3A46412AE1AF72AB66511433B80E18B00938E2626A82BC70CC05E
GY = 0x693F46716EB6BC248876203756C9C7624BEA73736CA3984087789C1E05A0C2D73AD3FF1CE67C39C4FDBD132C4ED7C8AD9808795BF230FA14
GENERATOR = (GX, GY)
IDENTITY = (0, 1)


def inv(a: int) -> int:
    return pow(a, P - 2, P)


def on_curve(pt) -> bool:
    x, y = pt
    return (x * x + y * y - 1 - D * x * x * y * y) % P == 0


# ---- extended coordinates (X:Y:Z:T), a = 1 ----------------------
def to_ext(pt):
    x, y = pt
    return (x % P, y % P, 1, x * y % P)


def to_affine(e):
    """ExtendedPoint::to_affine (call sites src/ecc/signable.rs:49,79)."""
    X, Y, Z, _ = e
    zi = inv(Z)
    return (X * zi % P, Y * zi % P)


def ext_add(p, q):
    """add-2008-hwcd, a = 1, complete because d is a non-square."""
    X1, Y1, Z1, T1 = p
    X2, Y2, Z2, T2 = q
    A = X1 * X2 % P
    B = Y1 * Y2 % P
    C = D * T1 % P * T2 % P
    Dd = Z1 * Z2 % P
    E = ((X1 + Y1) * (X2 + Y2) - A - B) % P
    F = (Dd - C) % P
    G = (Dd + C) % P
    H = (B - A) % P
    return (E * F % P, G * H % P, F * G % P, E * H % P)


def ext_double(p):
    return ext_add(p, p)


def ext_neg(p):
    X, Y, Z, T = p
    return ((-X) % P, Y, Z, (-T) % P)


def scalar_mult_ext(k: int, pt_ext):
    """`ExtendedPoint * Scalar` for the exact (possibly unreduced, up to 448-bit) integer k
    (quirk Q10: verify multiplies by the unreduced h, src/ecc/signable.rs:76-77)."""
    acc = to_ext(IDENTITY)
    for i in reversed(range(k.bit_length())):
        acc = ext_double(acc)
        if (k >> i) & 1:
            acc = ext_add(acc, pt_ext)
    return acc


def scalar_mult(k: int, pt):
    return to_affine(scalar_mult_ext(k, to_ext(pt)))


def point_add(p, q):
    return to_affine(ext_add(to_ext(p), to_ext(q)))


# ---- byte codecs ---------------------------------------------------------------------
def bytes_to_scalar(b: bytes) -> int:
    """src/sha3/aux_functions.rs:102-106 -- 56 bytes big-endian, NOT reduced."""
    if len(b) != 56:
        raise ValueError("BytesToScalarError")
    return int.from_bytes(b, "big")


def scalar_to_bytes(s: int) -> bytes:
    """src/sha3/aux_functions.rs:108-110."""
    return s.to_bytes(56, "big")


def fe_to_bytes(x: int) -> bytes:
    """FieldElement::to_bytes -- canonical, 56 bytes little-endian."""
    return (x % P).to_bytes(56, "little")


def point_to_bytes(pt) -> bytes:
    """Our parity format for a point: affine x || y, each 56-byte LE canonical."""
    return fe_to_bytes(pt[0]) + fe_to_bytes(pt[1])


def point_from_bytes(b: bytes):
    return (int.from_bytes(b[:56], "little"), int.from_bytes(b[56:112], "little"))


# ---- protocol -------------------------------------------------
def secret_scalar(pw: bytes, d: int) -> int:
    """s = 4 * BE(KMACXOF(pw, "", 448, "SK", d)) mod r  (src/ecc/keypair.rs:42-43)."""
    return bytes_to_scalar(ref_sha3.kmac_xof(pw, b"", 448, b"SK", d)) * 4 % R


def keygen(pw: bytes, d: int):
    """KeyPair::new public-key derivation V = [s]G (src/ecc/keypair.rs:41-51), affine."""
    return scalar_mult(secret_scalar(pw, d), GENERATOR)


def sign(pw: bytes, msg: bytes, d: int) -> tuple[bytes, bytes]:
    """Signable::sign (src/ecc/signable.rs:40-57).  Returns (h: 56 bytes, z: 56 bytes BE)."""
    s = secret_scalar(pw, d)
    s_bytes = scalar_to_bytes(s)
    k = bytes_to_scalar(ref_sha3.kmac_xof(s_bytes, msg, 448, b"N", d)) * 4 % R
    U = scalar_mult(k, GENERATOR)
    h = ref_sha3.kmac_xof(fe_to_bytes(U[0]), msg, 448, b"T", d)
    z = (k - bytes_to_scalar(h) * s % R) % R
    return h, scalar_to_bytes(z)


def verify(pub, msg: bytes, h: bytes, z_be: bytes, d: int) -> bool:
    """Signable::verify (src/ecc/signable.rs:72-86): U = [z]G + [BE(h)]V, h unreduced."""
    z = bytes_to_scalar(z_be)
    U = ext_add(scalar_mult_ext(z, to_ext(GENERATOR)), scalar_mult_ext(bytes_to_scalar(h), to_ext(pub)))
    h_p = ref_sha3.kmac_xof(fe_to_bytes(to_affine(U)[0]), msg, 448, b"T", d)
    return h_p == h


def key_encrypt(pub, msg: bytes, d: int, k_rand56: bytes):
    """KeyEncryptable::key_encrypt (src/ecc/encryptable.rs:34-50) with the 56 random bytes
    injected.  Returns (ciphertext, tag t, nonce point Z affine)."""
    k = bytes_to_scalar(k_rand56) * 4 % R
    W = scalar_mult(k, pub)
    Z = scalar_mult(k, GENERATOR)
    ke_ka = ref_sha3.kmac_xof(fe_to_bytes(W[0]), b"", 448 * 2, b"PK", d)
    ke, ka = ke_ka[:56], ke_ka[56:]
    t = ref_sha3.kmac_xof(ka, msg, 448, b"PKA", d)
    stream = ref_sha3.kmac_xof(ke, b"", len(msg) * 8, b"PKE", d)
    return bytes(a ^ b for a, b in zip(msg, stream)), t, Z


def key_decrypt(pw: bytes, ct: bytes, d: int, Z, tag: bytes):
    """KeyEncryptable::key_decrypt (src/ecc/encryptable.rs:72-94).  Returns (ok, buffer):
    on failure the buffer is restored to the ciphertext (:88-93)."""
    s = secret_scalar(pw, d)
    W = scalar_mult(s, Z)
    ke_ka = ref_sha3.kmac_xof(fe_to_bytes(W[0]), b"", 448 * 2, b"PK", d)
    ke, ka = ke_ka[:56], ke_ka[56:]
    stream = ref_sha3.kmac_xof(ke, b"", len(ct) * 8, b"PKE", d)
    pt = bytes(a ^ b for a, b in zip(ct, stream))
    t_p = ref_sha3.kmac_xof(ka, pt, 448, b"PKA", d)
    if t_p == tag:
        return True, pt
    return False, bytes(ct)


# ---- helpers for the independent OpenSSL cross-checks ---------------------------------
def rfc8032_encode(pt) -> bytes:
    """57-byte Ed448 point encoding (y LE, top bit of last byte = lsb of x)."""
    x, y = pt
    b = bytearray(y.to_bytes(57, "little"))
    b[56] |= (x & 1) << 7
    return bytes(b)


def montgomery_u(pt) -> int:
    """RFC 7748 4.2: edwards448 (x, y) -> curve448 u = y^2 / x^2."""
    x, y = pt
    return y * y % P * inv(x * x % P) % P

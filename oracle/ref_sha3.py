"""CPU oracle (TEST INFRASTRUCTURE ONLY) -- literal Python restatement of capyCRYPT's SHA3 path.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this
module.  The product path (capycrypt_b200/) never does.

Every function cites the reference file:line it follows (paths relative to
/root/reference).  The reference deviates from FIPS 202 / SP 800-185 in several corner
cases; "bit-exact" means reproducing those, so this file
restates the reference's *steps*, not the standards.

Parity status: PINNED -- tests/test_oracle_kat.py checks this module against every
known-answer vector the reference's own tests hold for the path
(src/sha3/shake_functions.rs:92-288, src/sha3/sponge.rs:99-190,
tests/integration_tests.rs:84-93), stored in tests/golden/sha3_kat.json.
"""
from __future__ import annotations

MASK64 = (1 << 64) - 1

# src/sha3/keccakf.rs:12-37
RC = [
    0x0000000000000001, 0x0000000000008082, 0x800000000000808A, 0x8000000080008000,
    0x000000000000808B, 0x0000000080000001, 0x8000000080008081, 0x8000000000008009,
    0x000000000000008A, 0x0000000000000088, 0x0000000080008009, 0x000000008000000A,
    0x000000008000808B, 0x800000000000008B, 0x8000000000008089, 0x8000000000008003,
    0x8000000000008002, 0x8000000000000080, 0x000000000000800A, 0x800000008000000A,
    0x8000000080008081, 0x8000000000008080, 0x0000000080000001, 0x8000000080008008,
]

# rho offsets for lane a[x + 5y] (FIPS 202 3.2.2); the reference hard-codes the same
# constants inside its unrolled rounds (src/sha3/keccakf.rs:66-146).
RHO = [
    0, 1, 62, 28, 27,
    36, 44, 6, 55, 20,
    3, 10, 43, 25, 39,
    41, 45, 15, 21, 8,
    18, 2, 61, 56, 14,
]


def _rol(x: int, n: int) -> int:
    n %= 64
    return ((x << n) | (x >> (64 - n))) & MASK64 if n else x


def keccakf_1600(a: list[int]) -> None:
    """Keccak-f[1600] on 25 u64 lanes a[x+5y], in place (src/sha3/keccakf.rs:8-423).

    The reference unrolls 4 rounds and works in place; the function computed is the
    standard 24-round permutation, written here round by round.
    """
    for rnd in range(24):
        c = [a[x] ^ a[x + 5] ^ a[x + 10] ^ a[x + 15] ^ a[x + 20] for x in range(5)]
        d = [c[(x + 4) % 5] ^ _rol(c[(x + 1) % 5], 1) for x in range(5)]
        b = [0] * 25
        for x in range(5):
            for y in range(5):
                # pi: B[y][2x+3y] = rot(A[x][y])
                b[y + 5 * ((2 * x + 3 * y) % 5)] = _rol(a[x + 5 * y] ^ d[x], RHO[x + 5 * y])
        for y in range(5):
            for x in range(5):
                a[x + 5 * y] = b[x + 5 * y] ^ ((~b[(x + 1) % 5 + 5 * y]) & MASK64 & b[(x + 2) % 5 + 5 * y])
        a[0] ^= RC[rnd]


# --------------------------------------------------------------------------------------
# src/sha3/constants.rs, src/lib.rs
# --------------------------------------------------------------------------------------
SEC_PARAMS = (224, 256, 384, 512)  # src/lib.rs:113-135
RATE_IN_BYTES = 136  # src/sha3/constants.rs:3


def capacity_from_bit_length(bit_length: int) -> int:
    """src/sha3/constants.rs:38-45."""
    x = bit_length * 2
    if x <= 448:
        return 448
    if x <= 512:
        return 512
    if x <= 768:
        return 768
    return 1024


def bytepad_value(d: int) -> int:
    """src/lib.rs:137-144."""
    return {224: 172, 256: 168, 384: 152, 512: 136}[d]


def check_d(d: int) -> int:
    """SecParam::try_from (src/lib.rs:126-134)."""
    if d not in SEC_PARAMS:
        raise ValueError("UnsupportedSecurityParameter")
    return d


# --------------------------------------------------------------------------------------
# src/sha3/sponge.rs
# --------------------------------------------------------------------------------------
def pad_ten_one(m: bytearray, rate_in_bytes: int) -> None:
    """src/sha3/sponge.rs:89-95 -- q zero bytes whose last is 0x80 (never ORed)."""
    q = rate_in_bytes - len(m) % rate_in_bytes
    padded = bytearray(q)
    padded[q - 1] = 0x80
    m += padded


def bytes_to_state(in_val: bytes, rate_in_bytes: int) -> list[int]:
    """src/sha3/sponge.rs:47-60.  Note len//rate blocks but (rate*8)//64 lanes per block:
    for rate 172 (D224 cSHAKE) the read offset advances 168 per block (quirk Q7)."""
    offset = 0
    s = [0] * 25
    lanes = (rate_in_bytes * 8) // 64
    for _ in range(len(in_val) // rate_in_bytes):
        for i in range(lanes):
            s[i] ^= int.from_bytes(in_val[offset:offset + 8], "little")  # bytes_to_word :63-69
            offset += 8
        keccakf_1600(s)
    return s


def sponge_absorb(m: bytearray, capacity_bits: int) -> list[int]:
    """src/sha3/sponge.rs:10-17 -- pads only when len % r != 0 (quirk Q1)."""
    r = (1600 - capacity_bits) // 8
    if len(m) % r != 0:
        pad_ten_one(m, r)
    return bytes_to_state(bytes(m), r)


def sponge_squeeze(s: list[int], bit_length: int, rate_bits: int) -> bytes:
    """src/sha3/sponge.rs:25-34 -- emits rate_bits//64 lanes per block and permutes after
    every block (the last permutation's result is dropped, quirk Q8)."""
    out = bytearray()
    block_size = rate_bits // 64
    while len(out) * 8 < bit_length:
        for v in s[0:block_size]:
            out += v.to_bytes(8, "little")  # state_to_byte_array :37-44
        keccakf_1600(s)
    del out[bit_length // 8:]
    return bytes(out)


# --------------------------------------------------------------------------------------
# src/sha3/aux_functions.rs (nist_800_185)
# --------------------------------------------------------------------------------------
def left_encode(value: int) -> bytes:
    """src/sha3/aux_functions.rs:34-49."""
    if value == 0:
        return bytes([1, 0])
    b = value.to_bytes(8, "big").lstrip(b"\x00")
    return bytes([len(b)]) + b


def right_encode(value: int) -> bytes:
    """src/sha3/aux_functions.rs:55-68 -- only correct for 0 (quirk Q6); restated as is."""
    if value == 0:
        return bytes([0, 1])
    b = bytearray(value.to_bytes(8, "big"))
    i = 1
    while i < 8 and b[i] == 0:
        i += 1
    b[0] = 9 - i
    return bytes(b[0:9 - i])


def encode_string(s: bytes) -> bytes:
    """src/sha3/aux_functions.rs:24-28."""
    return left_encode(len(s) * 8) + bytes(s)


def byte_pad(inp: bytes, w: int) -> bytes:
    """src/sha3/aux_functions.rs:11-18 -- always appends w - len%w zeros (a whole block of
    zeros when already aligned, quirk Q3)."""
    z = left_encode(w) + bytes(inp)
    padlen = w - (len(z) % w)
    return z + bytes(padlen)


# --------------------------------------------------------------------------------------
# src/sha3/shake_functions.rs
# --------------------------------------------------------------------------------------
def shake(n: bytearray, d: int) -> bytes:
    """SHA3-d (src/sha3/shake_functions.rs:24-32).  Mutates n exactly as the reference
    mutates Message.msg (quirk Q5).  Suffix choice uses rate 136 for every d (Q2)."""
    bytes_to_pad = RATE_IN_BYTES - len(n) % RATE_IN_BYTES
    n.append(0x86 if bytes_to_pad == 1 else 0x06)
    c = capacity_from_bit_length(d)
    s = sponge_absorb(n, c)
    return sponge_squeeze(s, d, 1600 - d)


def cshake(x: bytes, l: int, n: bytes, s: bytes, d: int) -> bytes:
    """src/sha3/shake_functions.rs:49-64.  Capacity = d bits (quirk Q7); when N and S are
    both empty the buffer is first mutated by shake() whose result is discarded (Q4)."""
    check_d(d)
    encoded = encode_string(n) + encode_string(s)
    out = bytearray(byte_pad(encoded, bytepad_value(d)))
    out += bytes(x)
    out.append(0x04)
    if len(n) == 0 and len(s) == 0:
        shake(out, d)
    st = sponge_absorb(out, d)
    return sponge_squeeze(st, l, 1600 - d)


def kmac_xof(k: bytes, x: bytes, l: int, s: bytes, d: int) -> bytes:
    """src/sha3/shake_functions.rs:79-89."""
    check_d(d)
    bp = bytearray(byte_pad(encode_string(k), bytepad_value(d)))
    bp += bytes(x)
    bp += right_encode(0)
    return cshake(bytes(bp), l, b"KMAC", s, d)


# --------------------------------------------------------------------------------------
# src/sha3/hashable.rs -- the drop-in operator surface
# --------------------------------------------------------------------------------------
def compute_sha3_hash(msg: bytes, d: int) -> bytes:
    """SpongeHashable::compute_sha3_hash (src/sha3/hashable.rs:19-21) on the ORIGINAL
    message bytes (the reference also leaves Message.msg suffixed+padded, Q5)."""
    check_d(d)
    return shake(bytearray(msg), d)


def compute_tagged_hash(msg: bytes, pw: bytes, s: bytes, d: int) -> bytes:
    """SpongeHashable::compute_tagged_hash (src/sha3/hashable.rs:33-35)."""
    return kmac_xof(pw, msg, d, s, d)


# --------------------------------------------------------------------------------------
# src/sha3/encryptable.rs -- sponge AE ("next" row N2); nonce z injected for determinism
# --------------------------------------------------------------------------------------
def sha3_encrypt(msg: bytes, pw: bytes, d: int, z: bytes) -> tuple[bytes, bytes]:
    """src/sha3/encryptable.rs:29-45 with the 512-byte nonce supplied by the caller.
    Returns (ciphertext, tag)."""
    ke_ka = kmac_xof(bytes(z) + bytes(pw), b"", 1024, b"S", d)
    ke, ka = ke_ka[:64], ke_ka[64:]
    t = kmac_xof(ka, msg, 512, b"SKA", d)
    m = kmac_xof(ke, b"", len(msg) * 8, b"SKE", d)
    return bytes(a ^ b for a, b in zip(msg, m)), t


def sha3_decrypt(ct: bytes, pw: bytes, d: int, z: bytes, tag: bytes) -> tuple[bool, bytes]:
    """src/sha3/encryptable.rs:58-83.  Returns (ok, message-buffer-after-call): on tag
    mismatch the buffer is restored to the ciphertext (:77-82)."""
    ke_ka = kmac_xof(bytes(z) + bytes(pw), b"", 1024, b"S", d)
    ke, ka = ke_ka[:64], ke_ka[64:]
    m = kmac_xof(ke, b"", len(ct) * 8, b"SKE", d)
    pt = bytes(a ^ b for a, b in zip(ct, m))
    new_t = kmac_xof(ka, pt, 512, b"SKA", d)
    if new_t == tag:
        return True, pt
    return False, bytes(ct)


# --------------------------------------------------------------------------------------
# "No reference counterpart": FIPS 202 SHAKE256, exposed by the GPU engine as an extra and
# checked against hashlib (BASELINE.json config 2 names SHAKE256; the reference has none).
# --------------------------------------------------------------------------------------
def fips_shake(msg: bytes, out_bytes: int, bits: int = 256) -> bytes:
    rate = (1600 - 2 * bits) // 8
    m = bytearray(msg)
    m.append(0x1F)
    m += bytes((-len(m)) % rate)
    m[-1] |= 0x80
    s = [0] * 25
    for o in range(0, len(m), rate):
        for i in range(rate // 8):
            s[i] ^= int.from_bytes(m[o + 8 * i:o + 8 * i + 8], "little")
        keccakf_1600(s)
    out = bytearray()
    while len(out) < out_bytes:
        for v in s[:rate // 8]:
            out += v.to_bytes(8, "little")
        if len(out) < out_bytes:
            keccakf_1600(s)
    return bytes(out[:out_bytes])

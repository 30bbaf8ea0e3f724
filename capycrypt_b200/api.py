"""Host-side mirror of capyCRYPT's operator surface for BATCHES (the `capycrypt::gpu` module of the
north star, in Python because the image has no Rust toolchain; the C++ twin is host/capycrypt_gpu.hpp).

Same names, argument meaning and error behaviour as the reference:
  SecParam                      src/lib.rs:113-135      (try_from -> UnsupportedSecurityParameter)
  Message, Signature            src/lib.rs:65-94, src/ecc/signable.rs:17-24
  KeyPair::new                  src/ecc/keypair.rs:41-51
  SpongeHashable                src/sha3/hashable.rs:7-36   compute_sha3_hash / compute_tagged_hash
  Signable                      src/ecc/signable.rs:12-15   sign / verify
  kmac_xof                      src/sha3/shake_functions.rs:79-89
  SpongeEncryptable             src/sha3/encryptable.rs:7-10  sha3_encrypt / sha3_decrypt
  KeyEncryptable                src/ecc/encryptable.rs:10-13  key_encrypt / key_decrypt
Every function forwards to the CUDA engine through the C ABI; nothing is computed on the CPU.
All operations work IN PLACE on the Message objects, like the reference.
"""
from __future__ import annotations

import datetime
import os
from dataclasses import dataclass, field
from enum import IntEnum
from typing import Optional, Sequence

import numpy as np

from .engine import Engine, pack


class OperationError(Exception):
    """src/lib.rs:9-30 -- the variant name is carried in `.kind`."""

    def __init__(self, kind: str):
        super().__init__(kind)
        self.kind = kind


class SecParam(IntEnum):
    D224 = 224
    D256 = 256
    D384 = 384
    D512 = 512

    @staticmethod
    def try_from(value: int) -> "SecParam":
        """src/lib.rs:126-134."""
        try:
            return SecParam(value)
        except ValueError:
            raise OperationError("UnsupportedSecurityParameter") from None

    def bit_length_(self) -> int:
        return int(self)


@dataclass
class Signature:
    """src/ecc/signable.rs:17-24: h = keyed hash (56 bytes), z = scalar (56 bytes big-endian, canonical)."""
    h: bytes
    z: bytes


@dataclass
class Message:
    """src/lib.rs:65-94."""
    msg: bytearray
    d: Optional[SecParam] = None
    sym_nonce: Optional[bytes] = None
    asym_nonce: Optional[bytes] = None  # affine point x || y (112 bytes)
    digest: bytes = b""
    sig: Optional[Signature] = None
    kem_ciphertext: Optional[bytes] = b""

    @staticmethod
    def new(data: bytes) -> "Message":
        return Message(msg=bytearray(data))


@dataclass
class KeyPair:
    """src/ecc/keypair.rs:13-22.  pub_key is the affine point x || y (112 bytes LE); priv_key is the password."""
    owner: str
    pub_key: bytes
    priv_key: bytes
    date_created: str = field(default_factory=lambda: datetime.datetime.now().strftime("%Y-%m-%d %H:%M:%S"))


class Gpu:
    """`capycrypt::gpu`: batch entry points over one engine context."""

    def __init__(self, engine: Optional[Engine] = None, devices=None):
        self.engine = engine or Engine(devices)

    # ---- SpongeHashable --------------------------------------------------------------------------
    def compute_sha3_hash(self, msgs: Sequence[Message], d: SecParam, mutate_like_reference: bool = False) -> None:
        """Batched Message::compute_sha3_hash: sets every `.digest`.  With mutate_like_reference the suffix and
        padding are also appended to `.msg` as the reference does (quirk Q5, shake_functions.rs:24-29)."""
        d = SecParam.try_from(int(d))
        data, off = pack([m.msg for m in msgs])
        out = self.engine.sha3(data, off, int(d))
        for m, row in zip(msgs, out):
            m.digest = row.tobytes()
            if mutate_like_reference:
                _append_reference_padding(m.msg, int(d))

    def compute_tagged_hash(self, msgs: Sequence[Message], pws: Sequence[bytes], s: str | bytes, d: SecParam) -> None:
        d = SecParam.try_from(int(d))
        s = s.encode() if isinstance(s, str) else s
        kd, ko = pack(pws)
        xd, xo = pack([m.msg for m in msgs])
        out = self.engine.kmac_xof(kd, ko, xd, xo, int(d), s, int(d))
        for m, row in zip(msgs, out):
            m.digest = row.tobytes()

    def kmac_xof(self, keys: Sequence[bytes], xs: Sequence[bytes], l: int, s: str | bytes, d: SecParam) -> list[bytes]:
        d = SecParam.try_from(int(d))
        s = s.encode() if isinstance(s, str) else s
        kd, ko = pack(keys)
        xd, xo = pack(xs)
        return [r.tobytes() for r in self.engine.kmac_xof(kd, ko, xd, xo, l, s, int(d))]

    # ---- KeyPair::new ----------------------------------------------------------------------------------
    def new_keypairs(self, pws: Sequence[bytes], owner: str, d: SecParam) -> list[KeyPair]:
        d = SecParam.try_from(int(d))
        pd, po = pack(pws)
        pub = self.engine.ed448_keygen(pd, po, int(d))
        return [KeyPair(owner=owner, pub_key=row.tobytes(), priv_key=bytes(pw)) for pw, row in zip(pws, pub)]

    # ---- Signable ----------------------------------------------------------------------------------------
    def sign(self, msgs: Sequence[Message], keys: Sequence[KeyPair], d: SecParam) -> None:
        d = SecParam.try_from(int(d))
        pd, po = pack([k.priv_key for k in keys])
        md, mo = pack([m.msg for m in msgs])
        h, z = self.engine.ed448_sign(pd, po, md, mo, int(d))
        for m, hr, zr in zip(msgs, h, z):
            m.sig = Signature(h=hr.tobytes(), z=zr.tobytes())
            m.d = d

    def verify(self, msgs: Sequence[Message], pub_keys: Sequence[bytes]) -> list[Optional[OperationError]]:
        """Per message: None for Ok(()), else the OperationError the reference would return
        (SignatureNotSet / SecurityParameterNotSet / SignatureVerificationFailure, signable.rs:72-86)."""
        res: list[Optional[OperationError]] = [None] * len(msgs)
        groups: dict[int, list[int]] = {}
        for i, m in enumerate(msgs):
            if m.sig is None:
                res[i] = OperationError("SignatureNotSet")
            elif m.d is None:
                res[i] = OperationError("SecurityParameterNotSet")
            elif len(m.sig.h) != 56 or len(m.sig.z) != 56 or len(pub_keys[i]) != 112:
                # a malformed signature or key fails on its own and is left out of the fixed-stride batch (one short
                # field would otherwise shift every later item); the reference rejects it at the type level
                res[i] = OperationError("SignatureVerificationFailure")
            else:
                groups.setdefault(int(m.d), []).append(i)
        for d, idx in groups.items():
            md, mo = pack([msgs[i].msg for i in idx])
            pub = np.frombuffer(b"".join(pub_keys[i] for i in idx), dtype=np.uint8)
            h = np.frombuffer(b"".join(msgs[i].sig.h for i in idx), dtype=np.uint8)
            z = np.frombuffer(b"".join(msgs[i].sig.z for i in idx), dtype=np.uint8)
            _, ok = self.engine.ed448_verify(pub, md, mo, h, z, d)
            for i, o in zip(idx, ok):
                if not o:
                    res[i] = OperationError("SignatureVerificationFailure")
        return res


    # ---- SpongeEncryptable ---------------------------------------------------------------------------
    def sha3_encrypt(self, msgs: Sequence[Message], pws: Sequence[bytes], d: SecParam,
                     nonces: Optional[Sequence[bytes]] = None) -> None:
        """Batched Message::sha3_encrypt (sha3/encryptable.rs:29-45): replaces .msg with the ciphertext, .digest with
        the tag, .sym_nonce with z, .d with d.  z = 512 random bytes per message (get_random_bytes(512), :31), drawn
        here on the host unless injected."""
        d = SecParam.try_from(int(d))
        if nonces is None:
            nonces = [os.urandom(512) for _ in msgs]
        if len(nonces) != len(msgs) or len(pws) != len(msgs) or any(len(z) != 512 for z in nonces):
            raise ValueError("sha3_encrypt: one password and one 512-byte nonce per message")
        pd, po = pack(pws)
        md, mo = pack([m.msg for m in msgs])
        ct, tag = self.engine.sponge_encrypt(pd, po, b"".join(nonces), 512, md, mo, int(d))
        for i, m in enumerate(msgs):
            m.msg = bytearray(ct[int(mo[i]):int(mo[i + 1])].tobytes())
            m.digest = tag[i].tobytes()
            m.sym_nonce = bytes(nonces[i])
            m.d = d

    def sha3_decrypt(self, msgs: Sequence[Message], pws: Sequence[bytes]) -> list[Optional[OperationError]]:
        """Batched Message::sha3_decrypt (:58-83).  Per message None for Ok(()), else SecurityParameterNotSet /
        SymNonceNotSet / SHA3DecryptionFailure; on failure .msg keeps the ciphertext."""
        res: list[Optional[OperationError]] = [None] * len(msgs)
        groups: dict[int, list[int]] = {}
        for i, m in enumerate(msgs):
            if m.d is None:
                res[i] = OperationError("SecurityParameterNotSet")
            elif m.sym_nonce is None:
                res[i] = OperationError("SymNonceNotSet")
            else:
                groups.setdefault((int(m.d), len(m.sym_nonce)), []).append(i)
        for (d, nl), idx in groups.items():
            pd, po = pack([pws[i] for i in idx])
            cd, co = pack([msgs[i].msg for i in idx])
            tags = b"".join(bytes(msgs[i].digest).ljust(64, b"\0")[:64] for i in idx)
            out, ok = self.engine.sponge_decrypt(pd, po, b"".join(msgs[i].sym_nonce for i in idx), nl, cd, co, tags, d)
            for j, i in enumerate(idx):
                good = bool(ok[j]) and len(msgs[i].digest) == 64
                if good:
                    msgs[i].msg = bytearray(out[int(co[j]):int(co[j + 1])].tobytes())
                else:
                    res[i] = OperationError("SHA3DecryptionFailure")
        return res

    # ---- KeyEncryptable ------------------------------------------------------------------------------
    def key_encrypt(self, msgs: Sequence[Message], pub_keys: Sequence[bytes], d: SecParam,
                    k_rand: Optional[Sequence[bytes]] = None) -> None:
        """Batched Message::key_encrypt (ecc/encryptable.rs:34-50): .msg <- ciphertext, .digest <- tag (56 bytes),
        .asym_nonce <- Z = [k]G (affine x || y), .d <- d.  k = 56 random bytes per message (:36) unless injected."""
        d = SecParam.try_from(int(d))
        if k_rand is None:
            k_rand = [os.urandom(56) for _ in msgs]
        if len(pub_keys) != len(msgs) or len(k_rand) != len(msgs):
            raise ValueError("key_encrypt: one public key and one nonce per message")
        for i in range(len(msgs)):
            if len(pub_keys[i]) != 112 or len(k_rand[i]) != 56:
                # (an ExtendedPoint / a 56-byte get_random_bytes in the reference: cannot be malformed there)
                raise ValueError(f"key_encrypt: item {i}: public key must be 112 bytes and the nonce 56 bytes")
        md, mo = pack([m.msg for m in msgs])
        rc, ct, tag, z = self.engine.ed448_key_encrypt(b"".join(pub_keys), b"".join(k_rand), md, mo, int(d))
        if rc:
            raise OperationError("InvalidPublicKey")
        for i, m in enumerate(msgs):
            m.msg = bytearray(ct[int(mo[i]):int(mo[i + 1])].tobytes())
            m.digest = tag[i].tobytes()
            m.asym_nonce = z[i].tobytes()
            m.d = d

    def key_decrypt(self, msgs: Sequence[Message], pws: Sequence[bytes]) -> list[Optional[OperationError]]:
        """Batched Message::key_decrypt (:72-94): None for Ok(()), else SymNonceNotSet (sic, :73) /
        SecurityParameterNotSet / KeyDecryptionError; on failure .msg keeps the ciphertext."""
        res: list[Optional[OperationError]] = [None] * len(msgs)
        groups: dict[int, list[int]] = {}
        for i, m in enumerate(msgs):
            if m.asym_nonce is None:
                res[i] = OperationError("SymNonceNotSet")
            elif m.d is None:
                res[i] = OperationError("SecurityParameterNotSet")
            elif len(m.asym_nonce) != 112:
                res[i] = OperationError("KeyDecryptionError")  # not a point encoding: fails alone, stays out of the batch
            else:
                groups.setdefault(int(m.d), []).append(i)
        for d, idx in groups.items():
            pd, po = pack([pws[i] for i in idx])
            cd, co = pack([msgs[i].msg for i in idx])
            tags = b"".join(bytes(msgs[i].digest).ljust(56, b"\0")[:56] for i in idx)
            _, out, ok = self.engine.ed448_key_decrypt(pd, po, b"".join(msgs[i].asym_nonce for i in idx), cd, co, tags, d)
            for j, i in enumerate(idx):
                good = bool(ok[j]) and len(msgs[i].digest) == 56
                if good:
                    msgs[i].msg = bytearray(out[int(co[j]):int(co[j + 1])].tobytes())
                else:
                    res[i] = OperationError("KeyDecryptionError")
        return res


def _append_reference_padding(msg: bytearray, d: int) -> None:
    """What shake() leaves behind in Message.msg (shake_functions.rs:24-29 + sponge.rs:13-15,89-95)."""
    msg.append(0x86 if len(msg) % 136 == 135 else 0x06)
    c = {224: 448, 256: 512, 384: 768, 512: 1024}[d]
    r = (1600 - c) // 8
    if len(msg) % r:
        q = r - len(msg) % r
        msg.extend(bytes(q - 1) + b"\x80")

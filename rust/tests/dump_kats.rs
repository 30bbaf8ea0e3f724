//! N1 hand-off kit: run this ONCE on a machine with cargo, inside a checkout of capyCRYPT 0.7.5
//! (so that the real `tiny_ed448_goldilocks 0.1.8` computes the values), and copy the JSON it writes back to
//! `tests/golden/ed448_crate_kat.json` of the B200 engine.  `tests/test_crate_kat.py` then pins the oracle AND the
//! engine to the crate: Ed448 parity turns from "unpinned" into "pinned".
//!
//!     cp <engine>/rust/tests/dump_kats.rs            <capycrypt>/tests/dump_kats.rs
//!     cp <engine>/tests/golden/ed448_kat_inputs.json <capycrypt>/tests/ed448_kat_inputs.json
//!     cd <capycrypt> && cargo test --test dump_kats -- --nocapture      # writes tests/ed448_crate_kat.json
//!     cp tests/ed448_crate_kat.json <engine>/tests/golden/
//!
//! SOURCE-ONLY: the engine's build image has no Rust toolchain, so this file has not been compiled.  It only uses
//! the reference's public API (`KeyPair::new`, `Signable::sign`, `kmac_xof`, the pub field `Scalar.val`) and the call sites the
//! reference itself has on the crate (`ExtendedPoint::generator() * s`, `to_affine()`, `.x.to_bytes()`), cited per line.
//!
//! Output schema = tests/golden/ed448_golden.json of the engine ("keygen", "sign", "key_encrypt", "sha3_encrypt" records
//! with hex fields) plus a "serde" section holding the crate's own JSON for a KeyPair, a signed Message and a Signature
//! (row N4: the wire formats).
use capycrypt::ecc::keypair::KeyPair;
use capycrypt::ecc::signable::Signable;
use capycrypt::sha3::shake_functions::kmac_xof;
use crypto_bigint::{Encoding, U448};
use capycrypt::{Message, SecParam};
use serde_json::{json, Value};
use tiny_ed448_goldilocks::curve::{extended_edwards::ExtendedPoint, field::scalar::Scalar};

/// `byte_utils::bytes_to_scalar` / `scalar_to_bytes` are pub(crate) (src/sha3/aux_functions.rs:102-110): same bodies
fn bytes_to_scalar(in_bytes: &[u8]) -> Scalar {
    Scalar { val: U448::from_be_slice(in_bytes) }
}
fn scalar_to_bytes(s: &Scalar) -> Vec<u8> {
    s.val.to_be_bytes().to_vec()
}
fn hex(b: &[u8]) -> String {
    b.iter().map(|x| format!("{:02x}", x)).collect()
}
fn unhex(s: &str) -> Vec<u8> {
    (0..s.len() / 2).map(|i| u8::from_str_radix(&s[2 * i..2 * i + 2], 16).unwrap()).collect()
}
/// affine x || y, 2 x 56 bytes little-endian: the point format of the engine's C ABI
fn xy(p: &ExtendedPoint) -> Vec<u8> {
    let a = p.to_affine();
    let mut o = a.x.to_bytes().to_vec();
    o.extend_from_slice(&a.y.to_bytes());
    o
}
fn xor(a: &[u8], b: &[u8]) -> Vec<u8> {
    a.iter().zip(b).map(|(x, y)| x ^ y).collect()
}

#[test]
fn dump_kats() {
    let inputs: Value = serde_json::from_str(&std::fs::read_to_string("tests/ed448_kat_inputs.json").unwrap()).unwrap();
    let (mut keygen, mut sign, mut key_encrypt, mut sha3_encrypt) = (vec![], vec![], vec![], vec![]);
    let mut serde_section = json!({});
    for (n, c) in inputs["cases"].as_array().unwrap().iter().enumerate() {
        let d_bits = c["d"].as_u64().unwrap() as usize;
        let d = SecParam::try_from(d_bits).unwrap();
        let pw = unhex(c["pw"].as_str().unwrap());
        let msg = unhex(c["msg"].as_str().unwrap());
        let k_rand = unhex(c["k_rand"].as_str().unwrap());
        let nonce = unhex(c["nonce"].as_str().unwrap());

        // KeyPair::new (src/ecc/keypair.rs:41-51)
        let kp = KeyPair::new(&pw, "kat".to_string(), d);
        keygen.push(json!({"d": d_bits, "pw": hex(&pw), "pub_xy": hex(&xy(&kp.pub_key))}));

        // Signable::sign (src/ecc/signable.rs:40-57) -- deterministic
        let mut m = Message::new(msg.clone());
        m.sign(&kp, d);
        let sig = m.sig.clone().unwrap();
        assert!(m.verify(&kp.pub_key).is_ok());
        sign.push(json!({"d": d_bits, "pw": hex(&pw), "msg": hex(&msg), "h": hex(&sig.h), "z": hex(&scalar_to_bytes(&sig.z))}));

        // KeyEncryptable::key_encrypt with an INJECTED nonce: the reference draws k from thread_rng
        // (src/ecc/encryptable.rs:36), so its steps are replayed here line by line with the crate's own operations
        let k = bytes_to_scalar(&k_rand).mul_mod(&Scalar::from(4_u64)); // :36
        let w = (kp.pub_key * k).to_affine(); // :37
        let z = (ExtendedPoint::generator() * k).to_affine(); // :38
        let ke_ka = kmac_xof(&w.x.to_bytes(), &[], 448 * 2, "PK", d); // :40
        let (ke, ka) = ke_ka.split_at(ke_ka.len() / 2); // :41
        let t = kmac_xof(ka, &msg, 448, "PKA", d); // :43
        let ks = kmac_xof(ke, &[], msg.len() * 8, "PKE", d); // :45
        let ct = xor(&msg, &ks);
        key_encrypt.push(json!({"d": d_bits, "pw": hex(&pw), "k_rand": hex(&k_rand), "msg": hex(&msg), "ct": hex(&ct),
                                "tag": hex(&t), "z_xy": hex(&xy(&z.to_extended()))}));

        // SpongeEncryptable::sha3_encrypt with an injected nonce (src/sha3/encryptable.rs:29-45)
        let mut ke_ka_in = nonce.clone(); // :33-34
        ke_ka_in.extend_from_slice(&pw);
        let ke_ka = kmac_xof(&ke_ka_in, &[], 1024, "S", d); // :36
        let (ke, ka) = ke_ka.split_at(64); // :37
        let t = kmac_xof(ka, &msg, 512, "SKA", d); // :39
        let ks = kmac_xof(ke, &[], msg.len() * 8, "SKE", d); // :41
        sha3_encrypt.push(json!({"d": d_bits, "pw": hex(&pw), "nonce": hex(&nonce), "msg": hex(&msg),
                                 "ct": hex(&xor(&msg, &ks)), "tag": hex(&t)}));

        if n == 5 {
            // row N4: the crate's own wire formats (serde JSON of KeyPair / Message / Signature)
            serde_section = json!({
                "keypair": serde_json::to_value(&kp).unwrap(),
                "message": serde_json::to_value(&m).unwrap(),
                "signature": serde_json::to_value(&sig).unwrap(),
                "pub_xy": hex(&xy(&kp.pub_key)), "h": hex(&sig.h), "z": hex(&scalar_to_bytes(&sig.z)),
            });
        }
    }
    // the generator question: [s]G for the RFC 8032 7.4 scalars, as affine x || y
    let rfc: Vec<Value> = inputs["rfc8032"].as_array().unwrap().iter().map(|v| {
        let s = bytes_to_scalar(&unhex(v["scalar_be56"].as_str().unwrap()));
        json!({"scalar_be56": v["scalar_be56"], "xy": hex(&xy(&(ExtendedPoint::generator() * s)))})
    }).collect();
    let out = json!({"source": "capycrypt 0.7.5 + tiny_ed448_goldilocks 0.1.8 (tests/dump_kats.rs)", "rfc8032_xy": rfc,
                     "keygen": keygen, "sign": sign, "key_encrypt": key_encrypt, "sha3_encrypt": sha3_encrypt,
                     "serde": serde_section});
    std::fs::write("tests/ed448_crate_kat.json", serde_json::to_string_pretty(&out).unwrap()).unwrap();
    println!("wrote tests/ed448_crate_kat.json");
}

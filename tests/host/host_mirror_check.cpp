// tests/host/host_mirror_check.cpp -- exercises the C++ host mirror (capycrypt_b200/host/capycrypt_gpu.hpp) the way
// the reference's own tests use its API (src/sha3/shake_functions.rs:92-203, tests/integration_tests.rs:63-81).
// Needs a GPU; run by tests/test_host_mirror_gpu.py.
#include <cstdio>
#include <cstring>
#include "../../capycrypt_b200/host/capycrypt_gpu.hpp"
using namespace capycrypt;

static std::string hex(const Bytes& b) {
  static const char* d = "0123456789abcdef";
  std::string s;
  for (uint8_t c : b) { s.push_back(d[c >> 4]); s.push_back(d[c & 15]); }
  return s;
}
#define EXPECT(c) do { if (!(c)) { printf("FAIL line %d: %s\n", __LINE__, #c); return 1; } } while (0)

int main() {
  gpu::Engine eng;
  // test_shake_256 / test_hashable: SHA3-256("") and SHA3-256("test")
  std::vector<Message> m;
  m.emplace_back(Bytes{});
  m.emplace_back(Bytes{'t', 'e', 's', 't'});
  EXPECT(!eng.compute_sha3_hash(m, 256));
  EXPECT(hex(m[0].digest) == "a7ffc6f8bf1ed76651c14756a061d662f580ff4de43b49fa82d80a4b80f8434a");
  EXPECT(hex(m[1].digest) == "36f028580bb02cc8272a9a020f4200e346e276ae664e45ee80745574e2f5ab80");
  // unsupported security parameter
  EXPECT(eng.compute_sha3_hash(m, 255) == OperationError::UnsupportedSecurityParameter);
  // test_compute_tagged_hash_512: pw = "test", s = "", msg = ""
  std::vector<Message> t;
  t.emplace_back(Bytes{});
  EXPECT(!eng.compute_tagged_hash(t, {Bytes{'t', 'e', 's', 't'}}, "", 512));
  EXPECT(hex(t[0].digest).substr(0, 16) == "0f9b5dcd47dc08e0");
  // test_kmac_256: KMACXOF256(K = 40..5f, X = 00010203, 64 bits, S = "My Tagged Application")
  Bytes key(32);
  for (int i = 0; i < 32; i++) key[i] = (uint8_t)(0x40 + i);
  std::vector<Message> x;
  x.emplace_back(Bytes{0, 1, 2, 3});
  auto out = eng.kmac_xof({key}, x, 64, "My Tagged Application", 512);
  EXPECT(hex(out[0]) == "1755133f1534752a");
  // test_sig_512: sign + verify round trip, then a wrong key and a tampered message must fail
  std::vector<Bytes> pws = {Bytes(64, 7), Bytes(16, 9)};
  auto keys = eng.new_keypairs(pws, "test key", 512);
  std::vector<Message> s;
  s.emplace_back(Bytes(5000, 0xab));
  s.emplace_back(Bytes(10, 0x11));
  eng.sign(s, keys, 512);
  auto r = eng.verify(s, {keys[0].pub_key, keys[1].pub_key});
  EXPECT(!r[0] && !r[1]);
  r = eng.verify(s, {keys[1].pub_key, keys[0].pub_key});
  EXPECT(r[0] == OperationError::SignatureVerificationFailure && r[1] == OperationError::SignatureVerificationFailure);
  s[0].msg[3] ^= 1;
  r = eng.verify(s, {keys[0].pub_key, keys[1].pub_key});
  EXPECT(r[0] == OperationError::SignatureVerificationFailure && !r[1]);
  Message unsigned_msg(Bytes{1, 2, 3});
  r = eng.verify({unsigned_msg}, {keys[0].pub_key});
  EXPECT(r[0] == OperationError::SignatureNotSet);
  // test_symmetric_encryptable (tests/integration_tests.rs:96-114, SHA3 half) + decryption_test (:250-262)
  {
    std::vector<Message> e;
    e.emplace_back(Bytes(5242, 0x5a));
    e.emplace_back(Bytes{});
    std::vector<Bytes> pw = {Bytes(64, 1), Bytes(3, 2)};
    const Bytes plain0 = e[0].msg;
    EXPECT(!eng.sha3_encrypt(e, pw, 512));
    EXPECT(e[0].msg != plain0 && e[0].msg.size() == plain0.size() && e[0].digest.size() == 64 && e[0].sym_nonce->size() == 512);
    const Bytes enc0 = e[0].msg;
    auto bad = eng.sha3_decrypt(e, {Bytes(64, 3), pw[1]});
    EXPECT(bad[0] == OperationError::SHA3DecryptionFailure && !bad[1] && e[0].msg == enc0);
    auto good = eng.sha3_decrypt(e, pw);
    EXPECT(!good[0] && !good[1] && e[0].msg == plain0);
    Message no_d(Bytes{1});
    std::vector<Message> nd = {no_d};
    EXPECT(eng.sha3_decrypt(nd, {Bytes{}})[0] == OperationError::SecurityParameterNotSet);
  }
  // test_key_gen_enc_dec_512 (:44-60) + test_key_decrypt_handling_bad_input (:268-281)
  {
    std::vector<Message> e;
    e.emplace_back(Bytes(5242, 0x33));
    e.emplace_back(Bytes(125, 0x44));
    const Bytes plain0 = e[0].msg, plain1 = e[1].msg;
    EXPECT(!eng.key_encrypt(e, {keys[0].pub_key, keys[1].pub_key}, 512));
    EXPECT(e[0].msg != plain0 && e[0].digest.size() == 56 && e[0].asym_nonce.has_value());
    const Bytes enc0 = e[0].msg, enc1 = e[1].msg;
    auto bad = eng.key_decrypt(e, {pws[1], pws[0]});
    EXPECT(bad[0] == OperationError::KeyDecryptionError && bad[1] == OperationError::KeyDecryptionError);
    EXPECT(e[0].msg == enc0 && e[1].msg == enc1);
    auto good = eng.key_decrypt(e, pws);
    EXPECT(!good[0] && !good[1] && e[0].msg == plain0 && e[1].msg == plain1);
  }
  printf("host mirror ok\n");
  return 0;
}

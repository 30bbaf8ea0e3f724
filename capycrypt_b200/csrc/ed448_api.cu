// placeholder until the Ed448 kernels land: every Ed448 entry point reports BAD_ARG
#include "internal.h"
namespace capy {
struct Ed448Tables {};
void ed448_tables_free(DeviceCtx& dc) { (void)dc; }
}  // namespace capy
extern "C" {
int capy_ed448_fixed_base_batch(capy_ctx*, const uint8_t*, uint64_t, uint8_t*) { return CAPY_ERR_BAD_ARG; }
int capy_ed448_fixed_base_batch_dev(capy_ctx*, int, void*, const uint8_t*, uint64_t, uint8_t*) { return CAPY_ERR_BAD_ARG; }
int capy_ed448_var_base_batch(capy_ctx*, const uint8_t*, const uint8_t*, uint64_t, uint8_t*) { return CAPY_ERR_BAD_ARG; }
int capy_ed448_var_base_batch_dev(capy_ctx*, int, void*, const uint8_t*, const uint8_t*, uint64_t, uint8_t*, int*) { return CAPY_ERR_BAD_ARG; }
int capy_ed448_keygen_batch(capy_ctx*, int, const uint8_t*, const uint64_t*, uint64_t, uint8_t*) { return CAPY_ERR_BAD_ARG; }
int capy_ed448_keygen_batch_dev(capy_ctx*, int, void*, int, const uint8_t*, const uint64_t*, uint64_t, uint8_t*) { return CAPY_ERR_BAD_ARG; }
int capy_ed448_sign_batch(capy_ctx*, int, const uint8_t*, const uint64_t*, const uint8_t*, const uint64_t*, uint64_t, uint8_t*, uint8_t*) { return CAPY_ERR_BAD_ARG; }
int capy_ed448_sign_batch_dev(capy_ctx*, int, void*, int, const uint8_t*, const uint64_t*, const uint8_t*, const uint64_t*, uint64_t, uint8_t*, uint8_t*) { return CAPY_ERR_BAD_ARG; }
int capy_ed448_verify_batch(capy_ctx*, int, const uint8_t*, const uint8_t*, const uint64_t*, const uint8_t*, const uint8_t*, uint64_t, uint8_t*) { return CAPY_ERR_BAD_ARG; }
int capy_ed448_verify_batch_dev(capy_ctx*, int, void*, int, const uint8_t*, const uint8_t*, const uint64_t*, const uint8_t*, const uint8_t*, uint64_t, uint8_t*, int*) { return CAPY_ERR_BAD_ARG; }
int capy_ed448_ecdh_batch(capy_ctx*, const uint8_t*, const uint8_t*, uint64_t, uint8_t*, uint8_t*) { return CAPY_ERR_BAD_ARG; }
}

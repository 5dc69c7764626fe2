// ssi_tc.cu — tensor-core (tcgen05/TMEM/TMA) path.  Placeholder until the kernel lands.
#include "ssi_common.cuh"

bool ssi_tc_supported(const ssi_ctx*) { return false; }
int ssi_tc_prepare(ssi_ctx*) { return SSI_OK; }
void ssi_tc_invalidate(ssi_ctx*) {}
void ssi_tc_destroy(ssi_ctx*) {}
int ssi_tc_sse(ssi_ctx* ctx, const float*, int64_t, double*) {
    return ssi_fail(ctx, SSI_ERR_UNSUPPORTED, "tensor path not built");
}

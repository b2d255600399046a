#include "common.cuh"
extern "C" int rlp_leduc_generate(int32_t, int32_t, int32_t, int32_t, int64_t *, int32_t *, int32_t *, int8_t *,
                                  int32_t *, double *, double *, double *, int32_t *, uint8_t *) {
    return rlp::fail(RLP_ERR_UNSUPPORTED, "leduc generator not built yet");
}

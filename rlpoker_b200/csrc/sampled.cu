#include "cfr.cuh"
extern "C" int rlp_cfr_sampled_iters(rlp_cfr *, int, int64_t, uint64_t, int64_t, int64_t, int, void *) {
    return rlp::fail(RLP_ERR_UNSUPPORTED, "sampled CFR not built yet");
}

// Tensor-core (tcgen05 / TMEM) tensor-product convolution -- placeholder entry points until the
// kernel lands (the fp32 path is the only conv path in this build).
#include "ddp_common.cuh"

extern "C" int64_t ddp_tpconv_pack_size(const ddp_tpconv_t *, int32_t) { return DDP_E_UNSUPPORTED; }
extern "C" int ddp_tpconv_pack(const ddp_tpconv_t *, const ddp_tp_group_t *, const float *, const float *, const float *,
                               const float *, int32_t, void *) { return DDP_E_UNSUPPORTED; }
extern "C" int ddp_tpconv_umma(const ddp_tpconv_t *, const void *, int32_t, const ddp_tpconv_edges_t *, float *, void *) {
    return DDP_E_UNSUPPORTED;
}

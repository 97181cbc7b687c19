// tn_render.cu -- fused forward render (placeholder until the fused kernels land in this round).
#include "tn_common.cuh"
namespace tn {
struct RenderState { int dummy; };
void free_render(tn_tracer *h) { delete h->render; h->render = nullptr; }
}  // namespace tn
extern "C" int tn_render_set_field(tn_tracer *, const float *, uint32_t, uint32_t, void *) { return tn::fail(TN_ERR_STATE, "tn_render: not built"); }
extern "C" int tn_render_set_weights(tn_tracer *, const float *const *, void *) { return tn::fail(TN_ERR_STATE, "tn_render: not built"); }
extern "C" int tn_render(tn_tracer *, const tn_render_config *, const float *, const float *, uint32_t, float *, float *, float *, uint8_t *, void *) {
    return tn::fail(TN_ERR_STATE, "tn_render: not built");
}

// TEMPORARY stubs for entry points whose implementation has not landed yet (removed as net.cu /
// ensemble.cu arrive). They fail loudly.
#include "kernels.h"
using namespace mgb;
#define STUB(name) { set_error(name ": not implemented in this build"); return MGB_ERR_UNSUPPORTED; }
extern "C" {
int mgb_create(const mgb_config*, mgb_handle**) STUB("mgb_create")
void mgb_destroy(mgb_handle*) {}
int mgb_load_tensor(mgb_handle*, const char*, const void*, const int64_t*, int32_t, int32_t) STUB("mgb_load_tensor")
int mgb_finalize_weights(mgb_handle*) STUB("mgb_finalize_weights")
int mgb_set_text_embedding(mgb_handle*, const float*, int32_t) STUB("mgb_set_text_embedding")
int mgb_set_schedule(mgb_handle*, int32_t, const int32_t*, const float*, const float*, const float*) STUB("mgb_set_schedule")
int mgb_encode(mgb_handle*, const float*, int32_t, int32_t, int32_t, float*, void*) STUB("mgb_encode")
int mgb_unet_step(mgb_handle*, const float*, float*, const float*, float*, int32_t, int32_t, int32_t, int32_t, void*) STUB("mgb_unet_step")
int mgb_denoise(mgb_handle*, const float*, float*, const float*, int32_t, int32_t, int32_t, void*) STUB("mgb_denoise")
int mgb_decode(mgb_handle*, const float*, int32_t, int32_t, int32_t, int32_t, float*, void*) STUB("mgb_decode")
int mgb_ens_depth_cost(mgb_handle*, const float*, const double*, int32_t, int64_t, int32_t, int32_t, int32_t, double, double*, void*) STUB("mgb_ens_depth_cost")
int mgb_ens_minmax(mgb_handle*, const float*, int32_t, int64_t, float*, float*, void*) STUB("mgb_ens_minmax")
int mgb_ens_depth_reduce(mgb_handle*, const float*, const double*, int32_t, int64_t, int32_t, int32_t, int32_t, float*, float*, int32_t*, void*) STUB("mgb_ens_depth_reduce")
int mgb_ens_normals(mgb_handle*, const float*, int32_t, int64_t, int32_t, float*, float*, int32_t*, void*) STUB("mgb_ens_normals")
size_t mgb_workspace_bytes(mgb_handle*, int32_t, int32_t, int32_t) { return 0; }
}

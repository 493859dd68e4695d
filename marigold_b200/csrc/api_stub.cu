// TEMPORARY stubs for entry points whose implementation has not landed yet (removed as net.cu /
// ensemble.cu arrive). They fail loudly.
#include "kernels.h"
using namespace mgb;
#define STUB(name) { set_error(name ": not implemented in this build"); return MGB_ERR_UNSUPPORTED; }
extern "C" {
int mgb_ens_depth_cost(mgb_handle*, const float*, const double*, int32_t, int64_t, int32_t, int32_t, int32_t, double, double*, void*) STUB("mgb_ens_depth_cost")
int mgb_ens_minmax(mgb_handle*, const float*, int32_t, int64_t, float*, float*, void*) STUB("mgb_ens_minmax")
int mgb_ens_depth_reduce(mgb_handle*, const float*, const double*, int32_t, int64_t, int32_t, int32_t, int32_t, float*, float*, int32_t*, void*) STUB("mgb_ens_depth_reduce")
int mgb_ens_normals(mgb_handle*, const float*, int32_t, int64_t, int32_t, float*, float*, int32_t*, void*) STUB("mgb_ens_normals")
}

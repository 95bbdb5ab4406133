// ndt.cu -- placeholder until the NDT kernels land (next commit): entry points fail loudly.
#include "common.cuh"
int ndt_align_device(rspcl_ctx* ctx, const rspcl_cloud*, const rspcl_cloud*, const rspcl_ndt_params*, const float*,
                     rspcl_ndt_result*, rspcl_cloud*) {
  RSPCL_FAIL(ctx, RSPCL_ERR_ARG, "NDT not built yet");
}
extern "C" void rspcl_ndt_reference_params(rspcl_ndt_params* p) {
  p->max_iterations = 50; p->min_points_per_voxel = 6; p->transformation_epsilon = 0.01; p->step_size = 0.1;
  p->outlier_ratio = 0.55; p->min_covar_eigvalue_mult = 0.01; p->resolution = 1.0f; p->_pad = 0;
}
extern "C" int rspcl_ndt_align(rspcl_ctx* ctx, const rspcl_cloud*, const rspcl_cloud*, const rspcl_ndt_params*, const float*,
                               rspcl_ndt_result*, rspcl_cloud*) { RSPCL_FAIL(ctx, RSPCL_ERR_ARG, "NDT not built yet"); }
extern "C" int rspcl_ndt_voxels(rspcl_ctx* ctx, const rspcl_cloud*, const rspcl_ndt_params*, void*, long long, int32_t*) {
  RSPCL_FAIL(ctx, RSPCL_ERR_ARG, "NDT not built yet"); }
extern "C" int rspcl_ndt_derivatives(rspcl_ctx* ctx, const rspcl_cloud*, const rspcl_cloud*, const rspcl_ndt_params*,
                                     const double*, double*, double*, double*) { RSPCL_FAIL(ctx, RSPCL_ERR_ARG, "NDT not built yet"); }

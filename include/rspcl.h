/*
 * rspcl.h -- C ABI of the B200-native (sm_100a) registration hot path of hyunminch/realsense-pointcloud.
 *
 * The reference has no FFI: its arithmetic boundary is the PCL C++ surface its scheme classes call
 * (SURVEY.md 8b).  Each entry point below names the reference call site(s) (relative to /root/reference/src)
 * and the PCL interface it replaces.  Plain pointers and sizes only; every function returns an int status
 * (RSPCL_OK = 0) and never throws; the message of the last failure is rspcl_last_error().  A context owns
 * one CUDA stream; handles are thread-compatible, not thread-safe.  There is no CPU fallback: without a
 * CUDA device rspcl_ctx_create fails with RSPCL_ERR_CUDA.
 *
 * Data model.  A `rspcl_cloud` is a BATCH of point clouds ("segments") resident in HBM: segment s holds
 * count[s] points at pts[s*stride .. s*stride+count[s]) with the counts kept on the device, so a chain of
 * operations (edges -> voxel filter -> ICP) never round-trips through the host.  A point is 16 bytes
 * {float x,y,z; uint32 rgba} (the .pcd "x y z rgb" row).  Host buffers may use that layout
 * (RSPCL_LAYOUT_PCD16) or pcl::PointXYZRGB's 32-byte in-memory layout (RSPCL_LAYOUT_PCL32:
 * float x,y,z,1; uint8 b,g,r,a; 12 bytes padding -- types.hpp:8).  4x4 transforms are column-major
 * float[16], i.e. Eigen::Matrix4f::data().
 */
#ifndef RSPCL_H
#define RSPCL_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RSPCL_OK 0
#define RSPCL_ERR_CUDA 1       /* CUDA runtime / launch failure, or no device */
#define RSPCL_ERR_ARG 2        /* invalid argument */
#define RSPCL_ERR_CAPACITY 3   /* output cloud stride too small */
#define RSPCL_ERR_RANGE 4      /* coordinates outside the grid key range */

#define RSPCL_LAYOUT_PCD16 0
#define RSPCL_LAYOUT_PCL32 1

typedef struct rspcl_ctx rspcl_ctx;
typedef struct rspcl_cloud rspcl_cloud;

/* ------------------------------------------------------------------ context */
int rspcl_ctx_create(int device, rspcl_ctx** out);
void rspcl_ctx_destroy(rspcl_ctx* ctx);
const char* rspcl_last_error(const rspcl_ctx* ctx);
int rspcl_ctx_sync(rspcl_ctx* ctx);
/* CUDA-event timer on the context's stream (the stream every kernel of this library is launched on). */
int rspcl_timer_start(rspcl_ctx* ctx);
int rspcl_timer_stop(rspcl_ctx* ctx, float* elapsed_ms);   /* synchronises */
/* Device time spanned by several contexts working concurrently (one stream each): the largest elapsed time between
 * any context's rspcl_timer_start event and any context's rspcl_timer_mark event.  Synchronises all of them. */
int rspcl_timer_mark(rspcl_ctx* ctx);
int rspcl_timer_span(rspcl_ctx* const* ctxs, int n, float* elapsed_ms);
/* number of kernels this library has launched on the context since creation */
long long rspcl_launch_count(const rspcl_ctx* ctx);
/* Per-kernel CUDA-event profile (off by default).  While enabled, the library brackets the launches of its main
 * kernels with events on the context stream and accumulates {elapsed ms, launches, work units}.  `kernel` is the
 * kernel name, e.g. "k_icp_step" (units = source points), "k_ndt_eval" (source points), "k_canny_nms" (pixels),
 * "k_transform2" (points), "k_approx_voxel" (points), "k_unpack" (points).  rspcl_profile_get synchronises. */
int rspcl_profile_enable(rspcl_ctx* ctx, int on);
int rspcl_profile_reset(rspcl_ctx* ctx);
int rspcl_profile_get(rspcl_ctx* ctx, const char* kernel, double* total_ms, long long* launches, double* units);
/* pinned host memory for e2e transfers */
int rspcl_host_alloc(rspcl_ctx* ctx, size_t bytes, void** out);
int rspcl_host_free(rspcl_ctx* ctx, void* p);

/* ------------------------------------------------------------------ clouds (types.hpp:8-10 rgb_point_cloud) */
int rspcl_cloud_create(rspcl_ctx* ctx, int n_seg, int stride, rspcl_cloud** out);
void rspcl_cloud_destroy(rspcl_ctx* ctx, rspcl_cloud* c);
int rspcl_cloud_n_seg(const rspcl_cloud* c);
int rspcl_cloud_stride(const rspcl_cloud* c);
/* organized dims (0,0 when unorganized) */
int rspcl_cloud_dims(const rspcl_cloud* c, int* width, int* height);
/* Host -> device.  `host` holds the segments back to back (sum(counts) points) in `layout`.  width*height
 * must equal every count for an organized batch, else pass 0,0.  Asynchronous on the context stream when
 * `host` is pinned (rspcl_host_alloc). */
int rspcl_cloud_upload(rspcl_ctx* ctx, rspcl_cloud* c, const void* host, int layout, const int32_t* counts,
                       int n_seg, int width, int height);
/* Device -> host: counts (n_seg ints) first; then the points, packed back to back, into `host`
 * (capacity_points points of `layout`).  Either pointer may be NULL.  Synchronises. */
int rspcl_cloud_counts(rspcl_ctx* ctx, const rspcl_cloud* c, int32_t* counts);
int rspcl_cloud_download(rspcl_ctx* ctx, const rspcl_cloud* c, void* host, int layout, long long capacity_points,
                         int32_t* counts);

/* The caller rewrote the colours of an organized batch through its own device code (or wants the gray plane of
 * edge_extractor.hpp:17-24's Canny input rebuilt): drop the cached (r+g+b)/3 plane; the next edge extraction recomputes it
 * from the points.  (bench.py uses it to keep the gray conversion INSIDE the timed device-resident step.) */
int rspcl_cloud_invalidate_gray(rspcl_ctx* ctx, rspcl_cloud* c);
/* Device -> host of the xyz part only, into a caller buffer that already holds pcl::PointXYZRGB points (32 B each) whose
 * colours are still valid -- pcl::transformPointCloud(in, out) on out == a copy of in changes 12 of the 32 bytes.  Writes
 * {x, y, z, 1.0f} (16 B) at a 32-byte pitch and leaves bytes 16..31 of every point untouched.  mode 0: packed staging +
 * 2-D DMA copy (cudaMemcpy2DAsync); mode 1: a kernel that stores straight into the PINNED host buffer (rspcl_host_alloc)
 * over PCIe.  Segments are packed back to back like rspcl_cloud_download.  Synchronises. */
int rspcl_cloud_download_xyz_pcl32(rspcl_ctx* ctx, const rspcl_cloud* c, void* host_pcl32, long long capacity_points,
                                   int mode);

/* blur_filter.hpp:18-36 BlurFilter::filter (centre 3/5 crop of an organized cloud; also capture.hpp:79-104) */
int rspcl_crop35(rspcl_ctx* ctx, const rspcl_cloud* in, rspcl_cloud* out);

/* edge_extractor.hpp:7-39 extract_edge_features -> pcl::OrganizedEdgeFromRGBNormals (RGB-Canny class,
 * label_indices[4]) + pcl::copyPointCloud(cloud, indices).  out_edges: one unorganized segment per frame, row-major
 * pixel order.  host_mask (optional, n_seg*w*h bytes, 255 = edge) forces a synchronising copy. */
int rspcl_edge_extract(rspcl_ctx* ctx, const rspcl_cloud* frames, float t_low, float t_high,
                       rspcl_cloud* out_edges, uint8_t* host_mask);

/* The per-pixel label image of pcl::OrganizedEdgeFromRGBNormals::compute as edge_extractor.hpp:17-24 configures it
 * (setDepthDisconThreshold(0.2), setMaxSearchNeighbors(50), all edge types), which rs-pcl --edges shows (main.cpp:58-74):
 * bit 1 NAN_BOUNDARY, 2 OCCLUDING, 4 OCCLUDED (OrganizedEdgeBase, depth only), 16 RGB_CANNY.  The HIGH_CURVATURE class
 * (bit 8: Canny on the integral-image normals, edge_extractor.hpp:10-15) is NOT computed.  host_labels: n_seg*w*h bytes. */
int rspcl_edge_labels(rspcl_ctx* ctx, const rspcl_cloud* frames, float th_depth_discon, int max_search_neighbors,
                      float t_low, float t_high, uint8_t* host_labels);

/* pcl::ApproximateVoxelGrid::setLeafSize/setInputCloud/filter (icp:47,59-60,75-76; ndt:45,57-58,68-69;
 * incr:54-55).  Order- and bit-exact restatement of the 512-slot streaming filter.  in == out is allowed. */
int rspcl_voxel_approx(rspcl_ctx* ctx, const rspcl_cloud* in, const float leaf[3], rspcl_cloud* out);
/* voxel coordinates (3 ints/point) and history slot of every input point, packed like a download */
int rspcl_voxel_keys(rspcl_ctx* ctx, const rspcl_cloud* in, const float leaf[3], int32_t* host_ijk,
                     int32_t* host_slot);

/* pcl::transformPointCloud(in, out, Matrix4f) (icp:116-117; ndt:104-105; incr:63).  T: n_seg matrices, or one
 * matrix applied to every segment when broadcast != 0.  in == out is allowed (icp:117). */
int rspcl_transform(rspcl_ctx* ctx, const rspcl_cloud* in, const float* T, int broadcast, rspcl_cloud* out);

/* pcl::PointCloud::operator+ (icp:57,119-120; ndt:55,107-108; incr:64): out[s] = a[s] then b[s]. */
int rspcl_concat(rspcl_ctx* ctx, const rspcl_cloud* a, const rspcl_cloud* b, rspcl_cloud* out);
/* copy segment src_seg of `src` into segment dst_seg of `dst` (device to device) */
int rspcl_cloud_copy_segment(rspcl_ctx* ctx, const rspcl_cloud* src, int src_seg, rspcl_cloud* dst, int dst_seg);

/* ------------------------------------------------------------------ ICP
 * pcl::IterativeClosestPoint<PointXYZRGB,PointXYZRGB>: setMaximumIterations / setMaxCorrespondenceDistance /
 * setTransformationEpsilon / setEuclideanFitnessEpsilon / setInputSource / setInputTarget / align(out[, guess]) /
 * hasConverged / getFinalTransformation (icp:35,41-52,78-79,95,104,108-113; ndt:32,47-50,96-101; incr:37,46-49,57-61).
 */
typedef struct rspcl_icp_params {
  int32_t max_iterations;            /* reference 100 */
  int32_t min_correspondences;       /* PCL 3 */
  double  max_corr_dist;             /* reference 0.01 */
  double  transformation_epsilon;    /* reference 1 */
  double  euclidean_fitness_epsilon; /* reference 1000 */
  double  mse_threshold_absolute;    /* PCL DefaultConvergenceCriteria 1e-12 */
} rspcl_icp_params;
void rspcl_icp_reference_params(rspcl_icp_params* p);   /* icp:42-45 */

enum { RSPCL_CONV_NOT_CONVERGED = 0, RSPCL_CONV_ITERATIONS = 1, RSPCL_CONV_TRANSFORM = 2, RSPCL_CONV_ABS_MSE = 3,
       RSPCL_CONV_REL_MSE = 4, RSPCL_CONV_NO_CORRESPONDENCES = 5 };

typedef struct rspcl_icp_result {
  float   T[16];        /* getFinalTransformation() */
  int32_t converged;    /* hasConverged() */
  int32_t state;        /* DefaultConvergenceCriteria::ConvergenceState */
  int32_t iterations;
  int32_t n_corr;       /* correspondences of the last executed iteration */
  double  mse;          /* their mean squared distance */
  double  prev_mse;     /* in/out: correspondences_prev_mse_ (persists across align() on one PCL object) */
} rspcl_icp_result;

/* Aligns segment s of `src` onto segment s of `tgt` for every s (independent pairs, one launch sequence).
 * guess: n_seg matrices or NULL (identity).  results: n_seg entries (prev_mse is read as input; set it to
 * DBL_MAX for a fresh object).  aligned (optional) receives final applied to the source, as align(out) does.
 * first_corr (optional host buffer, packed like a download of src): iteration-1 match index or -1.
 * tgt may have a single segment shared by all sources (n_seg_tgt == 1). */
int rspcl_icp_align(rspcl_ctx* ctx, const rspcl_cloud* src, const rspcl_cloud* tgt, const rspcl_icp_params* prm,
                    const float* guess, rspcl_icp_result* results, rspcl_cloud* aligned, int32_t* first_corr);

/* Debug / parity variant of rspcl_icp_align: additionally returns the correspondences of the first n_dump_iterations
 * iterations (pcl::registration::CorrespondenceEstimation::determineCorrespondences + the max-distance rejection, as
 * IterativeClosestPoint::computeTransformation runs them every iteration: icp:95,104,111).  host_corr holds
 * n_dump_iterations blocks, each packed like a download of src: index of the matched target point or -1.  Iterations the
 * align did not execute (earlier convergence) read -1. */
int rspcl_icp_align_dump(rspcl_ctx* ctx, const rspcl_cloud* src, const rspcl_cloud* tgt, const rspcl_icp_params* prm,
                         const float* guess, rspcl_icp_result* results, int n_dump_iterations, int32_t* host_corr);

/* Host-only helper (no device work): the per-pair cluster sizes the persistent ICP kernel would use for one wave of
 * pairs with `counts` source points, given `budget` SMs and the SMs a cluster of 1..8 CTAs occupies (`weights`, 8 entries,
 * NULL: the cluster size itself).  Returns the number of source points a CTA keeps register-resident. */
int rspcl_debug_plan_clusters(const int32_t* counts, int n, double budget, const double* weights, int32_t* out_cl);

/* pcl::Registration::getFitnessScore(max_range) on an already transformed source: mean squared NN distance over
 * the source points whose NN is within max_range (squared distance <= max_range), DBL_MAX if none. */
int rspcl_fitness(rspcl_ctx* ctx, const rspcl_cloud* src_transformed, const rspcl_cloud* tgt, double max_range,
                  double* fitness);
/* exact 1-NN of every source point (lowest-index tie-break): packed idx / squared distance (host buffers) */
int rspcl_nearest(rspcl_ctx* ctx, const rspcl_cloud* query, const rspcl_cloud* tgt, int32_t* host_idx,
                  float* host_d2);

/* ------------------------------------------------------------------ NDT
 * pcl::NormalDistributionsTransform<PointXYZRGB,PointXYZRGB>: setTransformationEpsilon / setStepSize /
 * setResolution / setMaximumIterations / setInputSource / setInputTarget / align(out, guess) /
 * getFinalTransformation (ndt:38-43,71-72,83,92,104).
 */
typedef struct rspcl_ndt_params {
  int32_t max_iterations;           /* reference 50 */
  int32_t min_points_per_voxel;     /* PCL 6 */
  double  transformation_epsilon;   /* reference 0.01 */
  double  step_size;                /* reference 0.1 */
  double  outlier_ratio;            /* PCL 0.55 */
  double  min_covar_eigvalue_mult;  /* PCL 0.01 */
  float   resolution;               /* reference 1.0 */
  int32_t _pad;
} rspcl_ndt_params;
void rspcl_ndt_reference_params(rspcl_ndt_params* p);   /* ndt:39-43 */

typedef struct rspcl_ndt_result {
  float   T[16];
  int32_t converged;
  int32_t iterations;
  int32_t n_derivative_evals;
  int32_t n_hessian_evals;
  double  trans_probability;
  double  score;
  double  p[6];
} rspcl_ndt_result;

int rspcl_ndt_align(rspcl_ctx* ctx, const rspcl_cloud* src, const rspcl_cloud* tgt, const rspcl_ndt_params* prm,
                    const float* guess, rspcl_ndt_result* results, rspcl_cloud* aligned);

/* voxel Gaussians of the target (VoxelGridCovariance::filter) for parity checks; record layout = 224 bytes:
 * int32 ijk[3], npts; float centroid[3], pad; double mean[3], cov[9], icov[9], evals[3].  Sorted by (seg,iz,iy,ix).
 * Returns the number of voxels per segment in n_vox (host, n_seg); capacity in records. */
int rspcl_ndt_voxels(rspcl_ctx* ctx, const rspcl_cloud* tgt, const rspcl_ndt_params* prm, void* host_records,
                     long long capacity, int32_t* n_vox);
/* one computeDerivatives evaluation at pose p[6] per segment (score, g[6], H[36] row-major, packed per segment) */
int rspcl_ndt_derivatives(rspcl_ctx* ctx, const rspcl_cloud* src, const rspcl_cloud* tgt,
                          const rspcl_ndt_params* prm, const double* p, double* score, double* g, double* H);

/* ------------------------------------------------------------------ point-sharded multi-GPU mode
 * One process per GPU.  Every rank passes its SHARD of the source points and a full replica of the target; the
 * per-iteration partial sums {n, sum s, sum t, sum s t^T, sum d^2} (17 fp64, ICP) / {score, g, H} (28 fp64, NDT) are
 * combined with an NCCL all-reduce on the context stream and every rank runs the identical solve, so all ranks return
 * the same transform / convergence state (n_corr is the global count).  `aligned` receives the rank's shard moved by
 * the final transform.  These calls are collective: every rank of the communicator must make them in the same order.
 * rspcl_comm_unique_id is called on one rank and the 128 bytes are distributed by the application (e.g.
 * torch.distributed broadcast, MPI). */
int rspcl_comm_unique_id(void* id128);
int rspcl_comm_init(rspcl_ctx* ctx, int nranks, int rank, const void* id128);
int rspcl_comm_destroy(rspcl_ctx* ctx);
int rspcl_icp_align_sharded(rspcl_ctx* ctx, const rspcl_cloud* src_shard, const rspcl_cloud* tgt, const rspcl_icp_params* prm,
                            const float* guess, rspcl_icp_result* results, rspcl_cloud* aligned);
int rspcl_ndt_align_sharded(rspcl_ctx* ctx, const rspcl_cloud* src_shard, const rspcl_cloud* tgt, const rspcl_ndt_params* prm,
                            const float* guess, rspcl_ndt_result* results, rspcl_cloud* aligned);

/* ------------------------------------------------------------------ pairwise registration pipeline
 * One call = the reference's per-frame body (icp:75-120 / ndt:68-108) for a batch of independent frame pairs
 * (SURVEY H5 pairwise formulation): edges of every frame once, 1 cm approximate voxel filter, coarse stage (ICP or
 * NDT) from `guess`, fine ICP from identity, and -- when out_transformed != NULL -- transformPointCloud of the
 * full source frame by both transforms.  Pair i registers frame src_idx[i] onto frame tgt_idx[i].
 */
enum { RSPCL_COARSE_ICP = 0, RSPCL_COARSE_NDT = 1 };
typedef struct rspcl_pair_result {
  float   T_coarse[16];
  float   T_fine[16];
  int32_t converged;        /* fine ICP hasConverged() (icp:113) */
  int32_t coarse_iterations;
  int32_t fine_iterations;
  int32_t n_corr;           /* fine stage, last iteration */
  int32_t n_src;            /* voxel-filtered source edge points */
  int32_t n_tgt;
  double  mse;              /* fine stage */
} rspcl_pair_result;

int rspcl_register_pairs(rspcl_ctx* ctx, const rspcl_cloud* frames, const int32_t* src_idx, const int32_t* tgt_idx,
                         int n_pairs, int coarse_kind, const rspcl_icp_params* icp, const rspcl_ndt_params* ndt,
                         const float leaf[3], float t_low, float t_high, const float* guess /* n_pairs x 16 */,
                         rspcl_pair_result* results, rspcl_cloud* out_transformed /* n_pairs segments or NULL */);

/* ------------------------------------------------------------------ sequential registration (the reference's own loop)
 * TwoPhaseRegistrationScheme::registration (types.hpp:30-43) + ICPEdgeBasedRegistration / NDTEdgeBasedRegistration ::
 * global_registration (icp:26-130 / ndt:23-117) for ONE sweep, device-resident from the uploaded frames to the merged
 * cloud: edges of every frame, 1 cm voxel filter, then for k = 1 .. n-1, IN ORDER: coarse stage (ICP or NDT) of frame
 * k's edges onto the ACCUMULATING edge target from guess[k], fine ICP from identity, and -- if the fine align converged
 * (icp:113) -- the full frame k moved by both transforms is appended to the merged cloud (icp:120) and the fine-aligned
 * edges are PREPENDED to the target (icp:119: new points first).  The target lives at the END of its buffer and grows
 * towards the front, so prepending costs the new points only (no O(total) copy per frame as in `*a = *b + *a`).
 * frames: organized batch, segment k = frame k.  guess: n x 16 (entry 0 unused) -- R_y(k * rads) for the fixed-angle
 * schemes (icp:98-100, ndt:86-88).  results: n entries (entry 0: converged = 1, identity transforms).
 * out_global: 1 segment of stride >= n * w * h.  out_target (optional): 1 segment of stride >= the sum of the voxel-
 * filtered edge counts -- the final edge target (what icp:126 dumps as dataset/edge_cloud.pcd). */
int rspcl_register_sequence(rspcl_ctx* ctx, const rspcl_cloud* frames, int coarse_kind, const rspcl_icp_params* icp,
                            const rspcl_ndt_params* ndt, const float leaf[3], float t_low, float t_high, const float* guess,
                            rspcl_pair_result* results, rspcl_cloud* out_global, rspcl_cloud* out_target);

#ifdef __cplusplus
}
#endif
#endif

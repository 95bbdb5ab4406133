/*
 * orc.h -- CPU ORACLE for the rs-pcl registration hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * This library is the checker, never the product: only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load it.  Nothing under
 * realsense-pointcloud_b200/ links, imports or calls it.
 *
 * PARITY UNPINNED: the reference (/root/reference) delegates every arithmetic step to PCL >= 1.9
 * (CMakeLists.txt:11), which is neither vendored nor installable offline, and ships no tests, golden
 * vectors or data.  This oracle restates the PCL 1.9.1 algorithms the reference's call sites reach
 *  and is pinned only by known-answer constants,
 * independent cross-checks (scipy cKDTree, numpy SVD/eigh, finite differences) and synthetic ground
 * truth.  Upstream PCL files each function follows are cited next to it.
 *
 * Conventions: a point is 16 bytes {x,y,z,rgba} (the .pcd "x y z rgb" row; rgba = a<<24|r<<16|g<<8|b
 * as in pcl::PointXYZRGB).  4x4 transforms are column-major float[16] (Eigen::Matrix4f::data()).
 */
#ifndef ORC_H
#define ORC_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct OrcPoint { float x, y, z; uint32_t rgba; } OrcPoint;

/* ---- edge extraction: edge_extractor.hpp:7-39 -> pcl::OrganizedEdgeFromRGB -> pcl::Edge::canny ---- */
void orc_gaussian_kernel3(float k[9]);                       /* pcl/2d/impl/kernel.hpp gaussianKernel (3, sigma 1) */
/* Optional debug planes may be NULL.  near_bin_edge counts pixels (mag>=lo) whose angle in degrees lies
 * within 1e-3 of a discretisation threshold (atan2f ulp hazard).  Returns #edge pixels. */
int orc_canny(const OrcPoint* cloud, int w, int h, float t_low, float t_high,
              uint8_t* mask, float* dbg_blur, float* dbg_gx, float* dbg_gy, float* dbg_mag,
              uint8_t* dbg_dir, float* dbg_maxima, int* near_bin_edge);
/* extract_edge_features: returns n, writes the RGB-Canny points (row-major order) and their pixel indices */
int orc_extract_edges(const OrcPoint* cloud, int w, int h, float t_low, float t_high,
                      OrcPoint* out, int32_t* out_idx);
/* pcl::OrganizedEdgeBase::extractEdges (features/impl/organized_edge_detection.hpp), as configured by
 * edge_extractor.hpp:19-21 (setDepthDisconThreshold(0.2), setMaxSearchNeighbors(50)): per pixel label bits
 * 1 = NAN_BOUNDARY, 2 = OCCLUDING, 4 = OCCLUDED from the depth (z) channel alone.  labels: w*h bytes. */
void orc_depth_edge_labels(const OrcPoint* cloud, int w, int h, float th_depth_discon, int max_search_neighbors,
                           uint8_t* labels);
/* blur_filter.hpp:18-36 (centre 3/5 crop).  out holds (w*3/5)*(h*3/5) points; returns that count. */
int orc_crop35(const OrcPoint* cloud, int w, int h, OrcPoint* out, int* out_w, int* out_h);

/* ---- pcl::ApproximateVoxelGrid::applyFilter (filters/impl/approximate_voxel_grid.hpp) ---- */
int orc_approx_voxel(const OrcPoint* in, int n, const float leaf[3], OrcPoint* out);
/* voxel coordinates + history slot per input point (for bit-exact key checks) */
void orc_voxel_keys(const OrcPoint* in, int n, const float leaf[3], int32_t* ijk, int32_t* slot);

/* ---- transforms (common/impl/transforms.hpp), concatenation ---- */
void orc_transform(const OrcPoint* in, int n, const float T[16], OrcPoint* out);

/* ---- exact 1-NN, squared L2 in float in FLANN L2_Simple order, lowest-index tie-break ---- */
void orc_nn_brute(const OrcPoint* tgt, int nt, const OrcPoint* q, int nq, int32_t* idx, float* d2);
void orc_nn_kdtree(const OrcPoint* tgt, int nt, const OrcPoint* q, int nq, int32_t* idx, float* d2);

/* ---- ICP: registration/impl/icp.hpp + default_convergence_criteria.hpp + umeyama ---- */
typedef struct OrcIcpParams {
  int    max_iterations;              /* Registration default 10; reference 100 (icp:42,49)            */
  double max_corr_dist;               /* reference 0.01 (icp:43,50)                                    */
  double transformation_epsilon;      /* reference 1 (icp:44,51)                                       */
  double euclidean_fitness_epsilon;   /* reference 1000 (icp:45,52)                                    */
  double mse_threshold_absolute;      /* DefaultConvergenceCriteria default 1e-12                      */
  int    min_correspondences;         /* 3                                                             */
  int    umeyama_float;               /* 1 = float Umeyama path (PCL literal), 0 = double "truth" path */
} OrcIcpParams;
void orc_icp_default_params(OrcIcpParams* p);      /* PCL Registration defaults */
void orc_icp_reference_params(OrcIcpParams* p);    /* the reference's literal settings */

enum { ORC_CONV_NOT_CONVERGED = 0, ORC_CONV_ITERATIONS = 1, ORC_CONV_TRANSFORM = 2, ORC_CONV_ABS_MSE = 3,
       ORC_CONV_REL_MSE = 4, ORC_CONV_NO_CORRESPONDENCES = 5 };

typedef struct OrcIcpResult {
  float  T[16];            /* final_transformation_ (column-major) */
  int    converged;
  int    state;
  int    iterations;
  int    n_corr;           /* correspondences of the last executed iteration */
  double mse;              /* mean correspondence squared distance, last iteration (before its update) */
  double prev_mse;         /* in/out: DefaultConvergenceCriteria::correspondences_prev_mse_ (persists across align) */
} OrcIcpResult;

/* aligned (may be NULL) receives final applied to the original source.  first_corr (may be NULL, ns ints)
 * receives the iteration-1 correspondence target index or -1. */
void orc_icp_align(const OrcPoint* src, int ns, const OrcPoint* tgt, int nt, const OrcIcpParams* prm,
                   const float guess[16], OrcIcpResult* res, OrcPoint* aligned, int32_t* first_corr);
/* the same align, returning the correspondences (target index or -1) of the first n_dump iterations: [n_dump][ns] */
void orc_icp_align_dump(const OrcPoint* src, int ns, const OrcPoint* tgt, int nt, const OrcIcpParams* prm,
                        const float guess[16], OrcIcpResult* res, int n_dump, int32_t* corr);
/* Umeyama on explicit pairs (pcl::umeyama, with_scaling=false). use_float selects Scalar. */
void orc_umeyama(const float* src_xyz, const float* tgt_xyz, int n, int use_float, float T[16]);
/* Registration::getFitnessScore on an already transformed source */
double orc_fitness(const OrcPoint* src_transformed, int ns, const OrcPoint* tgt, int nt, double max_range);

/* ---- NDT: registration/impl/ndt.hpp + filters/impl/voxel_grid_covariance.hpp ---- */
typedef struct OrcNdtParams {
  int    max_iterations;          /* PCL 35; reference 50 (ndt:43)  */
  double transformation_epsilon;  /* PCL 0.1; reference 0.01 (ndt:39) */
  double step_size;               /* 0.1 (ndt:40)  */
  float  resolution;              /* 1.0 (ndt:41)  */
  double outlier_ratio;           /* 0.55 */
  int    min_points_per_voxel;    /* 6 */
  double min_covar_eigvalue_mult; /* 0.01 */
} OrcNdtParams;
void orc_ndt_reference_params(OrcNdtParams* p);
void orc_ndt_gauss_constants(float resolution, double outlier_ratio, double* d1, double* d2);

typedef struct OrcNdtVoxel {
  int32_t ijk[3];      /* floor(p * inv_leaf) lattice coordinates */
  int32_t npts;        /* >= min_points, or -1 if rejected by the eigenvalue / inverse test */
  float   centroid[3]; /* float centroid used for the radius search */
  double  mean[3];
  double  cov[9];      /* after eigenvalue inflation; row-major */
  double  icov[9];
  double  evals[3];    /* raw eigenvalues ascending (before inflation) */
} OrcNdtVoxel;

typedef struct OrcNdtGrid OrcNdtGrid;
OrcNdtGrid* orc_ndt_grid_build(const OrcPoint* tgt, int nt, const OrcNdtParams* prm);
void orc_ndt_grid_free(OrcNdtGrid* g);
int  orc_ndt_grid_size(const OrcNdtGrid* g);                 /* voxels with >= min_points (incl. rejected) */
void orc_ndt_grid_get(const OrcNdtGrid* g, OrcNdtVoxel* out); /* sorted by (iz,iy,ix) leaf index as std::map */

/* computeDerivatives at pose p (tx,ty,tz,rx,ry,rz): source is transformed by the float matrix built from p
 * exactly as computeStepLengthMT does.  Returns score; g[6], H[36] row-major. n_pairs = #(point,voxel) pairs. */
double orc_ndt_derivatives(const OrcNdtGrid* grid, const OrcPoint* src, int ns, const OrcNdtParams* prm,
                           const double p[6], double g[6], double H[36], int compute_hessian, long long* n_pairs);

typedef struct OrcNdtResult {
  float  T[16];
  int    converged;
  int    iterations;
  double trans_probability;
  double score;
  double p[6];
  int    n_derivative_evals;
  int    n_hessian_evals;
} OrcNdtResult;
void orc_ndt_align(const OrcPoint* src, int ns, const OrcPoint* tgt, int nt, const OrcNdtParams* prm,
                   const float guess[16], OrcNdtResult* res, OrcPoint* aligned);
void orc_pose_to_matrix(const double p[6], float T[16]);  /* Translation * Rx * Ry * Rz in float */
void orc_matrix_to_pose(const float T[16], double p[6]);  /* translation + eulerAngles(0,1,2) (Eigen 3.3) */

/* ---- scheme drivers ---- */
typedef struct OrcSchemeStats {
  int    n_frames;
  int    n_accepted;
  double t_edges, t_voxel, t_coarse, t_fine, t_transform, t_total; /* seconds */
} OrcSchemeStats;

enum { ORC_COARSE_ICP = 0, ORC_COARSE_NDT = 1 };

/*
 * Pairwise registration of frame k onto frame k-1 (SURVEY H5 formulation used by the batched GPU path):
 *   edges(k-1), edges(k) -> voxel both -> coarse (ICP or NDT, guess) -> fine ICP (identity) -> T = T_fine*T_coarse.
 * Mirrors icp:59-60,75-111 / ndt:57-58,68-99 on one pair.  transformed_full (may be NULL) receives the
 * source frame transformed by both (icp:116-117).  Returns 1 if fine ICP converged.
 */
int orc_register_pair(const OrcPoint* frame_tgt, const OrcPoint* frame_src, int w, int h,
                      int coarse_kind, const OrcIcpParams* icp, const OrcNdtParams* ndt, const float leaf[3],
                      const float guess[16], float T_coarse[16], float T_fine[16],
                      OrcIcpResult* coarse_icp_res, OrcNdtResult* coarse_ndt_res, OrcIcpResult* fine_res,
                      OrcPoint* transformed_full, OrcSchemeStats* stats);

/*
 * Sequential schemes with a growing target, line by line after icp:26-130, ndt:23-117 (coarse_kind) with the
 * fixed-angle guess R_y(acc_rads) (use_imu=0) or IMU thetas (use_imu=1, thetas = 3*n floats, rewritten in place
 * like icp:84).  frames: n organized clouds of w*h points, concatenated.  out_global must hold n*w*h points;
 * returns its size.  T_out (may be NULL): per frame k>=1 the product T_fine*T_coarse (16 floats each), accepted[k].
 */
int orc_scheme_edge(const OrcPoint* frames, int n, int w, int h, int coarse_kind, int use_imu, float rads,
                    float* thetas, const OrcIcpParams* icp, const OrcNdtParams* ndt, const float leaf[3],
                    OrcPoint* out_global, float* T_out, int32_t* accepted, OrcSchemeStats* stats);
/* incremental_icp.hpp:35-69 (leaf: PCL default unless overridden).  out must hold n*w*h points. */
int orc_scheme_incremental(const OrcPoint* frames, int n, int npts, const OrcIcpParams* icp, const float leaf[3],
                           OrcPoint* out_target, float* T_out, int32_t* accepted);

#ifdef __cplusplus
}
#endif
#endif

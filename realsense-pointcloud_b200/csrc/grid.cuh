// grid.cuh -- uniform-grid / voxel-hash nearest-neighbour structure shared by ICP, fitness and NDT.
//
// Replaces the FLANN kd-tree PCL builds inside Registration::align (initCompute) for every setInputTarget
// (icp:79,109; ndt:72,97; incr:58).  A target batch is binned into cubic cells; occupied cells live in one
// open-addressing hash table keyed by (segment, ix, iy, iz) and point at a contiguous slice of a cell-sorted copy of
// the target points (counting sort by hash slot: count -> exclusive scan -> scatter).
#pragma once
#include "common.cuh"

struct DevGrid {
  unsigned long long* keys = nullptr;  // [cap] cell key or EMPTY
  int* cnt = nullptr;                  // [cap] points in the cell
  int* start = nullptr;                // [cap] first index into sorted
  float4* sorted = nullptr;            // [n_total] {x,y,z, original index within its segment (int bits)}
  unsigned cap_mask = 0;
  float inv_cs = 0.f;  // 1 / cell size
  float cs = 0.f;
  int shared_target = 0;  // 1: all queries use segment 0 of the target
  long long n_total = 0;
  int* slot_of = nullptr;  // [n_total] scratch (slot, rank) per target point
  int* rank_of = nullptr;
};

#define GRID_EMPTY 0xFFFFFFFFFFFFFFFFull

__device__ __forceinline__ unsigned long long grid_key(int seg, int ix, int iy, int iz) {
  return ((unsigned long long)(unsigned)seg << 48) | ((unsigned long long)(unsigned)(ix + 32768) << 32) |
         ((unsigned long long)(unsigned)(iy + 32768) << 16) | (unsigned long long)(unsigned)(iz + 32768);
}
__device__ __forceinline__ bool grid_in_range(int ix, int iy, int iz) {
  return ix > -32768 && ix < 32767 && iy > -32768 && iy < 32767 && iz > -32768 && iz < 32767;
}
__device__ __forceinline__ unsigned grid_hash(unsigned long long k) {
  k ^= k >> 33;
  k *= 0xff51afd7ed558ccdull;
  k ^= k >> 33;
  k *= 0xc4ceb9fe1a85ec53ull;
  k ^= k >> 33;
  return (unsigned)k;
}
__device__ __forceinline__ int grid_cell(float v, float inv_cs) { return __float2int_rd(__fmul_rn(v, inv_cs)); }

// slot of an occupied cell or -1
__device__ __forceinline__ int grid_lookup(const DevGrid& g, unsigned long long key) {
  unsigned s = grid_hash(key) & g.cap_mask;
  while (true) {
    unsigned long long k = __ldg(&g.keys[s]);
    if (k == key) return (int)s;
    if (k == GRID_EMPTY) return -1;
    s = (s + 1) & g.cap_mask;
  }
}

// Exact nearest neighbour among target points within `r` of q (r <= ~cs/2): scans the <= 2x2x2 cells that the ball
// touches.  Squared distance in FLANN L2_Simple order, ties -> lowest original index.  Returns -1 if none.
__device__ __forceinline__ int grid_nn_bounded(const DevGrid& g, int seg, float qx, float qy, float qz, float r,
                                               float* out_d2, float4* out_pt) {
  const int x0 = grid_cell(qx - r, g.inv_cs), x1 = grid_cell(qx + r, g.inv_cs);
  const int y0 = grid_cell(qy - r, g.inv_cs), y1 = grid_cell(qy + r, g.inv_cs);
  const int z0 = grid_cell(qz - r, g.inv_cs), z1 = grid_cell(qz + r, g.inv_cs);
  int best = -1;
  float bd = INFINITY;
  float4 bp = make_float4(0.f, 0.f, 0.f, 0.f);
  if (!grid_in_range(x0, y0, z0) || !grid_in_range(x1, y1, z1)) {
    *out_d2 = bd;
    return -1;
  }
  const int tseg = g.shared_target ? 0 : seg;
  for (int iz = z0; iz <= z1; ++iz)
    for (int iy = y0; iy <= y1; ++iy)
      for (int ix = x0; ix <= x1; ++ix) {
        const int s = grid_lookup(g, grid_key(tseg, ix, iy, iz));
        if (s < 0) continue;
        const int b = __ldg(&g.start[s]), e = b + __ldg(&g.cnt[s]);
        for (int k = b; k < e; ++k) {
          const float4 t = __ldg(&g.sorted[k]);
          const float d = dist2_l2simple(qx, qy, qz, t.x, t.y, t.z);
          const int idx = __float_as_int(t.w);
          if (d < bd || (d == bd && idx < best)) {
            bd = d;
            best = idx;
            bp = t;
          }
        }
      }
  *out_d2 = bd;
  if (out_pt) *out_pt = bp;
  return best;
}

// host API (grid.cu)
int grid_build(rspcl_ctx* ctx, const rspcl_cloud* tgt, float cell_size, DevGrid* g, int* d_range_flag);
void grid_free(rspcl_ctx* ctx, DevGrid* g);
// brute-force exact NN (unbounded): idx/d2 strided like the query cloud
int nn_brute_device(rspcl_ctx* ctx, const rspcl_cloud* query, const rspcl_cloud* tgt, int* d_idx, float* d_d2);

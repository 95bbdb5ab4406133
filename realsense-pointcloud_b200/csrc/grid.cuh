// grid.cuh -- uniform-grid / voxel-hash nearest-neighbour structure shared by ICP, fitness and NDT.
//
// Replaces the FLANN kd-tree PCL builds inside Registration::align (initCompute) for every setInputTarget
// (icp:79,109; ndt:72,97; incr:58).  A target batch is binned into cubic cells; occupied cells live in one
// open-addressing hash table keyed by (segment, ix, iy, iz).  A slot is 16 bytes {key, start, count} so that one
// 128-bit load resolves a probe, and points at a contiguous slice of a cell-sorted copy of the target points
// (counting sort by hash slot: count -> exclusive scan -> scatter).
#pragma once
#include "common.cuh"

struct __align__(16) GridSlot {
  unsigned long long key;  // cell key or GRID_EMPTY
  int start;               // first index into sorted
  int cnt;                 // points in the cell
};

struct DevGrid {
  GridSlot* slots = nullptr;  // [cap]
  float4* sorted = nullptr;   // [n_total] {x,y,z, original index within its segment (int bits)}
  unsigned cap_mask = 0;
  float inv_cs = 0.f;  // 1 / cell size
  float cs = 0.f;
  int shared_target = 0;  // 1: all queries use segment 0 of the target
  long long n_total = 0;
  // build-time scratch
  int* cnt = nullptr;      // [cap]
  int* start = nullptr;    // [cap]
  int* slot_of = nullptr;  // [n_total]
  int* rank_of = nullptr;  // [n_total]
};

#define GRID_EMPTY 0xFFFFFFFFFFFFFFFFull

__device__ __forceinline__ unsigned long long grid_key(int seg, int ix, int iy, int iz) {
  return ((unsigned long long)(unsigned)seg << 48) | ((unsigned long long)(unsigned)(ix + 32768) << 32) |
         ((unsigned long long)(unsigned)(iy + 32768) << 16) | (unsigned long long)(unsigned)(iz + 32768);
}
__device__ __forceinline__ bool grid_in_range(int ix, int iy, int iz) {
  return ix > -32768 && ix < 32767 && iy > -32768 && iy < 32767 && iz > -32768 && iz < 32767;
}
// 32-bit multiplicative hash of the cell coordinates (a handful of IMADs; the 64-bit key is what is compared)
__device__ __forceinline__ unsigned grid_hash4(int seg, int ix, int iy, int iz) {
  unsigned h = (unsigned)ix * 0x9E3779B1u ^ (unsigned)iy * 0x85EBCA77u ^ (unsigned)iz * 0xC2B2AE3Du ^ (unsigned)seg * 0x27D4EB2Fu;
  h ^= h >> 15;
  h *= 0x2C1B3C6Du;
  h ^= h >> 12;
  return h;
}
// generic 64-bit mix (NDT voxel table)
__device__ __forceinline__ unsigned grid_hash(unsigned long long k) {
  k ^= k >> 33;
  k *= 0xff51afd7ed558ccdull;
  k ^= k >> 33;
  k *= 0xc4ceb9fe1a85ec53ull;
  k ^= k >> 33;
  return (unsigned)k;
}
__device__ __forceinline__ int grid_cell(float v, float inv_cs) { return __float2int_rd(__fmul_rn(v, inv_cs)); }

__device__ __forceinline__ GridSlot grid_load_slot(const GridSlot* p) {
  const uint4 v = __ldg(reinterpret_cast<const uint4*>(p));
  GridSlot s;
  s.key = ((unsigned long long)v.y << 32) | v.x;
  s.start = (int)v.z;
  s.cnt = (int)v.w;
  return s;
}

// resolve a probe that did not hit on its first slot (rare: linear probing)
static __device__ __noinline__ GridSlot grid_probe_slow(const DevGrid& g, unsigned long long key, unsigned s) {
  while (true) {
    s = (s + 1) & g.cap_mask;
    GridSlot sl = grid_load_slot(&g.slots[s]);
    if (sl.key == key) return sl;
    if (sl.key == GRID_EMPTY) {
      sl.cnt = 0;
      return sl;
    }
  }
}

// Exact nearest neighbour among target points within `r` of q: scans the <= 2x2x2 cells that the ball touches
// (cell size >= 2r).  All eight first-probe loads are issued before any is consumed (memory-level parallelism);
// squared distance in FLANN L2_Simple order; ties -> lowest original index.  Returns -1 if none.
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
  GridSlot sl[8];
  unsigned long long keys[8];
  unsigned hs[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    // duplicate cells (when the ball does not straddle a cell face) are marked by an impossible key
    const bool dup = ((c & 1) && x1 == x0) || ((c & 2) && y1 == y0) || ((c & 4) && z1 == z0);
    const int ix = (c & 1) ? x1 : x0, iy = (c & 2) ? y1 : y0, iz = (c & 4) ? z1 : z0;
    keys[c] = dup ? GRID_EMPTY : grid_key(tseg, ix, iy, iz);
    hs[c] = grid_hash4(tseg, ix, iy, iz) & g.cap_mask;
  }
#pragma unroll
  for (int c = 0; c < 8; ++c) sl[c] = grid_load_slot(&g.slots[hs[c]]);
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    if (keys[c] == GRID_EMPTY) continue;
    GridSlot s = sl[c];
    if (s.key != keys[c]) {
      if (s.key == GRID_EMPTY) continue;
      s = grid_probe_slow(g, keys[c], hs[c]);
      if (s.cnt == 0) continue;
    }
    const int e = s.start + s.cnt;
    for (int k = s.start; k < e; ++k) {
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

// Same scan, additionally returning the position of the winner inside g.sorted and the smallest squared distance among
// the OTHER scanned points (d2nd), skipping position `skip` (the caller's cached incumbent, merged by the caller).
// Used by the certified-cache ICP passes (icp.cu): min(sqrt(d2nd), r) bounds the distance to every non-winner.
__device__ __forceinline__ int grid_nn_top2(const DevGrid& g, int seg, float qx, float qy, float qz, float r, int skip,
                                            float* out_d2, float* out_d2nd, int* out_pos, float4* out_pt) {
  const int x0 = grid_cell(qx - r, g.inv_cs), x1 = grid_cell(qx + r, g.inv_cs);
  const int y0 = grid_cell(qy - r, g.inv_cs), y1 = grid_cell(qy + r, g.inv_cs);
  const int z0 = grid_cell(qz - r, g.inv_cs), z1 = grid_cell(qz + r, g.inv_cs);
  int best = -1, bpos = -1;
  float bd = INFINITY, d2nd = INFINITY;
  float4 bp = make_float4(0.f, 0.f, 0.f, 0.f);
  if (grid_in_range(x0, y0, z0) && grid_in_range(x1, y1, z1)) {
    const int tseg = g.shared_target ? 0 : seg;
    // two halves of four cells (z0 then z1): half the slot / key registers of an eight-wide probe -- the callers are
    // latency-bound and run at four CTAs per SM, where registers are what limits the loads in flight
#pragma unroll 1
    for (int half = 0; half < 2; ++half) {
    if (half && z1 == z0) break;
    const int iz = half ? z1 : z0;
    GridSlot sl[4];
    unsigned long long keys[4];
    unsigned hs[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const bool dup = ((c & 1) && x1 == x0) || ((c & 2) && y1 == y0);
      const int ix = (c & 1) ? x1 : x0, iy = (c & 2) ? y1 : y0;
      keys[c] = dup ? GRID_EMPTY : grid_key(tseg, ix, iy, iz);
      hs[c] = grid_hash4(tseg, ix, iy, iz) & g.cap_mask;
    }
#pragma unroll
    for (int c = 0; c < 4; ++c) sl[c] = grid_load_slot(&g.slots[hs[c]]);
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      if (keys[c] == GRID_EMPTY) continue;
      GridSlot s = sl[c];
      if (s.key != keys[c]) {
        if (s.key == GRID_EMPTY) continue;
        s = grid_probe_slow(g, keys[c], hs[c]);
        if (s.cnt == 0) continue;
      }
      const int e = s.start + s.cnt;
      for (int k = s.start; k < e; ++k) {
        if (k == skip) continue;
        const float4 t = __ldg(&g.sorted[k]);
        const float d = dist2_l2simple(qx, qy, qz, t.x, t.y, t.z);
        const int idx = __float_as_int(t.w);
        if (d < bd || (d == bd && idx < best)) {
          d2nd = bd;
          bd = d;
          best = idx;
          bpos = k;
          bp = t;
        } else {
          d2nd = fminf(d2nd, d);
        }
      }
    }
    }
  }
  *out_d2 = bd;
  *out_d2nd = d2nd;
  *out_pos = bpos;
  *out_pt = bp;
  return best;
}

// Wider certificate for a query with nothing inside its gate ball: scan the 3x3x3 cells around the query's own cell.
// Every target point outside that block is at least one cell size away, so min(nearest found, cell size) bounds the
// distance to every target point and the query does not have to be looked at again until it has moved that far.
// Returns the nearest point found (position in g.sorted, -1 if none) and its squared distance / the runner-up's.
static __device__ __noinline__ int grid_scan27_top2(const DevGrid& g, int seg, float qx, float qy, float qz, float* out_d2,
                                             float* out_d2nd, float4* out_pt) {
  const int cx = grid_cell(qx, g.inv_cs), cy = grid_cell(qy, g.inv_cs), cz = grid_cell(qz, g.inv_cs);
  int bpos = -1, best = -1;
  float bd = INFINITY, d2nd = INFINITY;
  float4 bp = make_float4(0.f, 0.f, 0.f, 0.f);
  if (grid_in_range(cx - 1, cy - 1, cz - 1) && grid_in_range(cx + 1, cy + 1, cz + 1)) {
    const int tseg = g.shared_target ? 0 : seg;
#pragma unroll 1
    for (int c = 0; c < 27; ++c) {
      const int ix = cx + (c % 3) - 1, iy = cy + ((c / 3) % 3) - 1, iz = cz + (c / 9) - 1;
      const unsigned long long key = grid_key(tseg, ix, iy, iz);
      unsigned h = grid_hash4(tseg, ix, iy, iz) & g.cap_mask;
      GridSlot s = grid_load_slot(&g.slots[h]);
      if (s.key != key) {
        if (s.key == GRID_EMPTY) continue;
        s = grid_probe_slow(g, key, h);
        if (s.cnt == 0) continue;
      }
      const int e = s.start + s.cnt;
      for (int k = s.start; k < e; ++k) {
        const float4 t = __ldg(&g.sorted[k]);
        const float d = dist2_l2simple(qx, qy, qz, t.x, t.y, t.z);
        const int idx = __float_as_int(t.w);
        if (d < bd || (d == bd && idx < best)) {
          d2nd = bd;
          bd = d;
          best = idx;
          bpos = k;
          bp = t;
        } else {
          d2nd = fminf(d2nd, d);
        }
      }
    }
  } else {
    bd = -1.f;  // out of the key range: no certificate
  }
  *out_d2 = bd;
  *out_d2nd = d2nd;
  *out_pt = bp;
  return bpos;
}

// host API (grid.cu)
int grid_build(rspcl_ctx* ctx, const rspcl_cloud* tgt, float cell_size, DevGrid* g, int* d_range_flag);
void grid_free(rspcl_ctx* ctx, DevGrid* g);
// brute-force exact NN (unbounded): idx/d2 strided like the query cloud
int nn_brute_device(rspcl_ctx* ctx, const rspcl_cloud* query, const rspcl_cloud* tgt, int* d_idx, float* d_d2);

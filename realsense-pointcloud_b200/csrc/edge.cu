// edge.cu -- RGB-Canny edge extraction of organized clouds and ordered compaction into edge clouds.
//
// Replaces /root/reference/src/edge_extractor.hpp:7-39 (extract_edge_features).  Only label_indices[4]
// (EDGELABEL_RGB_CANNY) reaches the caller (edge_extractor.hpp:36-38) and it depends on rgb alone, so the
// integral-image normals and the four depth/curvature edge classes the reference computes and discards
// (edge_extractor.hpp:10-15,26-35) are not computed here.
//
// Kernels (all bit-exact against oracle/orc_edge.cpp; mul/add are explicitly un-fused):
//   K1  k_canny_nms   gray -> 3x3 Gaussian (9-tap, clamp) -> Sobel (clamp) -> sqrtf magnitude -> direction bin
//                     -> non-maximum suppression, fused over a shared-memory halo tile; writes a 1-byte class
//                     (0 none, 1 weak >= t_low, 2 strong >= t_high) and seeds the union-find parents.
//   K1b (in k_canny_nms) + k_uf_border / k_uf_flag / k_root_pull / k_edge_mask   hysteresis as 8-connected components of
//                     {class>0} that contain a strong pixel: lock-free union-find, tile-local in shared memory inside K1,
//                     stitched across tile borders in global memory over the TILE ROOTS only (the result is
//                     order-independent, so it is exact); the roots carry the strong flags in their class byte.
//   K1c k_block_count / k_seg_scan / k_scatter  mask -> ascending row-major compaction (copyPointCloud(indices)).
// Roofline: HBM-bound; algorithmic bytes per pixel = 1 R (gray) + 1 W (class) for K1.
#include "common.cuh"
#include <stdlib.h>

namespace {

constexpr int TW = 64, TH = 16;          // output tile
constexpr int GW = TW + 6, GH = TH + 6;  // gray tile (halo 3)
constexpr int BW = TW + 4, BH = TH + 4;  // blur tile (halo 2)
constexpr int MW = TW + 2, MH = TH + 2;  // magnitude tile (halo 1)

// pcl/2d/impl/kernel.hpp gaussianKernel(3, sigma=1) in float (checked by tests/test_oracle.py)
__constant__ float c_gauss[9] = {0.07511360943317413f, 0.12384141236543655f, 0.07511360943317413f,
                                 0.12384141236543655f, 0.20417995750904083f, 0.12384141236543655f,
                                 0.07511360943317413f, 0.12384141236543655f, 0.07511360943317413f};

__device__ __forceinline__ int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

// discretizeAngles on atan2f(gy,gx) * 57.29578f.  The angle is evaluated in double and rounded to float, which
// reproduces a correctly rounded atan2f except within ~1e-9 relative of a rounding boundary; the oracle counts pixels
// whose angle sits within 1e-3 degrees of a bin threshold (near_bin_edge) so a disagreement would be visible.
// (Slow path: only gradients within 1e-5 relative of a bin boundary get here, see direction_bin_fast.)
__device__ __noinline__ int direction_bin(float gy, float gx) {
  float rad = (float)atan2((double)gy, (double)gx);
  float angle = fmul(rad, 57.29578f);
  if (((angle <= 22.5f) && (angle >= -22.5f)) || (angle >= 157.5f) || (angle <= -157.5f)) return 0;
  if (((angle > 22.5f) && (angle < 67.5f)) || ((angle < -112.5f) && (angle > -157.5f))) return 45;
  if (((angle >= 67.5f) && (angle <= 112.5f)) || ((angle <= -67.5f) && (angle >= -112.5f))) return 90;
  if (((angle > 112.5f) && (angle < 157.5f)) || ((angle < -22.5f) && (angle > -67.5f))) return 135;
  return 255;
}

// The same bins from comparisons: the four thresholds +-22.5 / +-67.5 / +-112.5 / +-157.5 degrees are the lines
// |gy| = tan(22.5 deg) |gx| and |gy| = tan(67.5 deg) |gx|, and the diagonal bins differ by the sign of gy * gx.  A
// gradient within 1e-5 (relative) of a boundary -- where the rounding of atan2f and of the degree conversion, and the
// <= / < conventions of discretizeAngles, decide -- takes the atan2 path, so the result is the oracle's bin always.
__device__ __forceinline__ int direction_bin_fast(float gy, float gx) {
  const float a = fabsf(gy), b = fabsf(gx);
  const float l1 = 0.41421356f * b, l2 = 2.41421356f * b;
  if (fabsf(a - l1) <= 1e-5f * (a + l1) || fabsf(a - l2) <= 1e-5f * (a + l2)) return direction_bin(gy, gx);
  if (a < l1) return 0;
  if (a > l2) return 90;
  return ((gy > 0.f) == (gx > 0.f)) ? 45 : 135;
}

// class byte written by k_canny_nms: bits 0-1 = 0 none / 1 weak / 2 strong; CLS_ROOT = the pixel is the root of its tile-local
// component; CLS_LSTRONG (roots) = that local component holds a strong pixel; CLS_GSTRONG (roots, set by k_root_pull) = its
// whole 8-connected component does
constexpr unsigned CLS_LSTRONG = 4u, CLS_ROOT = 8u, CLS_GSTRONG = 16u;

// ---- lock-free union-find over candidate pixels (per frame; parents are pixel indices within the frame)
__device__ __forceinline__ int uf_find(volatile int* parent, int x) {
  int p = parent[x];
  while (p != x) {
    int gp = parent[p];
    if (gp != p) parent[x] = gp;  // path halving: gp is an ancestor, benign under concurrency
    x = p;
    p = gp;
  }
  return x;
}

__device__ __forceinline__ void uf_union(int* parent, int a, int b) {
  while (true) {
    a = uf_find(parent, a);
    b = uf_find(parent, b);
    if (a == b) return;
    if (a < b) {
      int t = a;
      a = b;
      b = t;
    }
    int old = atomicCAS(&parent[a], a, b);  // link the larger root under the smaller
    if (old == a) return;
  }
}

// steps 2 and 3 of k_canny_nms.  INTERIOR = the tile and its 3-pixel halo lie inside the image (three tiles in four at
// 640x480): every clamp is the identity and every position is in the image, so the index arithmetic collapses.
template <bool INTERIOR>
__device__ __forceinline__ void canny_blur_sobel(float (*s_gray)[GW + 1], float (*s_blur)[BW + 1], float (*s_mag)[MW + 1],
                                                 uint8_t (*s_dir)[TW], int tid, int r0, int c0, int w, int h, float t_low) {
  // 2. blur evaluated AT THE CLAMPED COORDINATE of every halo-2 position (what the Sobel pass will read)
  for (int k = tid; k < BH * BW; k += 256) {
    int i = k / BW, j = k % BW;
    int gi = i + 1, gj = j + 1;
    if (!INTERIOR) {
      int r = clampi(r0 - 2 + i, 0, h - 1), c = clampi(c0 - 2 + j, 0, w - 1);
      gi = r - (r0 - 3), gj = c - (c0 - 3);
    }
    float acc = 0.f;
#pragma unroll
    for (int kr = 0; kr < 3; ++kr)
#pragma unroll
      for (int kc = 0; kc < 3; ++kc) {
        // in-image neighbour clamp is already baked into s_gray as long as (r,c) itself is in the image
        acc = fadd(acc, fmul(c_gauss[kr * 3 + kc], s_gray[gi + kr - 1][gj + kc - 1]));
      }
    s_blur[i][j] = acc;
  }
  __syncthreads();
  // 3. Sobel + magnitude on the halo-1 region (only in-image positions are ever consumed)
  for (int k = tid; k < MH * MW; k += 256) {
    int i = k / MW, j = k % MW;
    int r = r0 - 1 + i, c = c0 - 1 + j;
    float m = 0.f;
    const bool centre = i >= 1 && i <= TH && j >= 1 && j <= TW;
    if (INTERIOR || (r >= 0 && r < h && c >= 0 && c < w)) {
      // blur(clamp(r+dr), clamp(c+dc)) lives at tile index (clamped coordinate - (origin-2))
      int bi[3], bj[3];
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        bi[d] = INTERIOR ? i + d : clampi(r + d - 1, 0, h - 1) - (r0 - 2);
        bj[d] = INTERIOR ? j + d : clampi(c + d - 1, 0, w - 1) - (c0 - 2);
      }
      // Sobel X {-1,0,1,-2,0,2,-1,0,1} and Y {-1,-2,-1,0,0,0,1,2,1} as the 9-tap correlations PCL runs, taps in row-major
      // order from a zero accumulator.  The zero taps add +-0 to a finite sum (the blurred image is finite) and the +-1 /
      // +-2 products are exact, so they are written as the adds and doublings they amount to: same bits, a third of
      // the operations.
      const float v00 = s_blur[bi[0]][bj[0]], v01 = s_blur[bi[0]][bj[1]], v02 = s_blur[bi[0]][bj[2]];
      const float v10 = s_blur[bi[1]][bj[0]], v12 = s_blur[bi[1]][bj[2]];
      const float v20 = s_blur[bi[2]][bj[0]], v21 = s_blur[bi[2]][bj[1]], v22 = s_blur[bi[2]][bj[2]];
      float gx = fadd(0.f, -v00);
      gx = fadd(gx, v02);
      gx = fadd(gx, -fadd(v10, v10));
      gx = fadd(gx, fadd(v12, v12));
      gx = fadd(gx, -v20);
      gx = fadd(gx, v22);
      float gy = fadd(0.f, -v00);
      gy = fadd(gy, -fadd(v01, v01));
      gy = fadd(gy, -v02);
      gy = fadd(gy, v20);
      gy = fadd(gy, fadd(v21, v21));
      gy = fadd(gy, v22);
      m = __fsqrt_rn(fadd(fmul(gx, gx), fmul(gy, gy)));
      // the direction is only ever read for centre pixels that pass the low threshold (suppressNonMaxima)
      if (centre) s_dir[i - 1][j - 1] = !(m < t_low) ? (uint8_t)direction_bin_fast(gy, gx) : (uint8_t)255;
    } else if (centre) {
      s_dir[i - 1][j - 1] = 255;
    }
    s_mag[i][j] = m;
  }
  __syncthreads();
}

// FROM_PTS: the frames carry no valid gray plane (their colours were rewritten on the device): the tile takes (r + g + b) / 3
// straight from the points instead of a separate 320 MB pass that would write the plane first.
template <bool FROM_PTS>
__global__ void __launch_bounds__(256) k_canny_nms(const uint8_t* __restrict__ gray, const float4* __restrict__ pts, int w, int h,
                                                   int stride, float t_low, float t_high, uint8_t* __restrict__ cls,
                                                   int* __restrict__ parent) {
  __shared__ float s_gray[GH][GW + 1];
  __shared__ float s_blur[BH][BW + 1];
  __shared__ float s_mag[MH][MW + 1];
  __shared__ uint8_t s_dir[TH][TW];  // direction bin of the centre pixels that can survive (magnitude >= t_low), else 255
  __shared__ uint8_t s_cls[TH][TW];
  __shared__ int s_par[TH * TW];  // tile-local union-find (parents are local pixel indices i * TW + j)
  __shared__ unsigned short s_list[TH * TW];  // the tile's candidates (any order)
  __shared__ uint8_t s_strong[TH * TW];       // per tile root: its local component holds a strong pixel
  __shared__ int s_ncand;
  const int seg = blockIdx.z;
  const int c0 = blockIdx.x * TW, r0 = blockIdx.y * TH;
  const int tid = threadIdx.y * blockDim.x + threadIdx.x;
  const uint8_t* g = FROM_PTS ? nullptr : gray + (size_t)seg * stride;
  if (tid == 0) s_ncand = 0;

  // 1. gray tile with replicate borders (Convolution BOUNDARY_OPTION_CLAMP).  Tiles whose halo lies inside the row (all
  //    but the first / last tile column) and whose rows are 16-byte aligned take 128-bit loads: the GW = 70 bytes of a
  //    tile row sit inside six aligned 16-byte words starting 16 bytes left of the tile.
  const bool vec_ok = !FROM_PTS && (w % 16 == 0) && (stride % 16 == 0) && c0 >= 16 && c0 + TW + 16 <= w;
  if (FROM_PTS) {
    const float4* P = pts + (size_t)seg * stride;
    for (int k = tid; k < GH * GW; k += 256) {
      int i = k / GW, j = k % GW;
      int r = clampi(r0 - 3 + i, 0, h - 1), c = clampi(c0 - 3 + j, 0, w - 1);
      const unsigned rgb = __float_as_uint(__ldg(&P[(size_t)r * w + c].w));
      s_gray[i][j] = (float)((((rgb >> 16) & 255u) + ((rgb >> 8) & 255u) + (rgb & 255u)) / 3u);  // == k_gray_from_pts
    }
  } else if (vec_ok) {
    for (int k = tid; k < GH * 6; k += 256) {
      const int i = k / 6, q = k % 6;
      const int r = clampi(r0 - 3 + i, 0, h - 1);
      const uint4 v = __ldg(reinterpret_cast<const uint4*>(g + (size_t)r * w + (c0 - 16)) + q);
      const unsigned wd[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int t = 0; t < 16; ++t) {
        const int j = 16 * q - 13 + t;  // tile column of byte t of word q (tile column 0 = image column c0 - 3)
        if (j >= 0 && j < GW) s_gray[i][j] = (float)((wd[t >> 2] >> (8 * (t & 3))) & 255u);
      }
    }
  } else {
    for (int k = tid; k < GH * GW; k += 256) {
      int i = k / GW, j = k % GW;
      int r = clampi(r0 - 3 + i, 0, h - 1), c = clampi(c0 - 3 + j, 0, w - 1);
      s_gray[i][j] = (float)g[(size_t)r * w + c];
    }
  }
  __syncthreads();
  // 2 + 3. blur and Sobel / magnitude / direction (see canny_blur_sobel)
  if (r0 >= 3 && r0 + TH + 3 <= h && c0 >= 3 && c0 + TW + 3 <= w)
    canny_blur_sobel<true>(s_gray, s_blur, s_mag, s_dir, tid, r0, c0, w, h, t_low);
  else
    canny_blur_sobel<false>(s_gray, s_blur, s_mag, s_dir, tid, r0, c0, w, h, t_low);
  // 4. direction + non-maximum suppression for the TH x TW centre (suppressNonMaxima: interior pixels only); the
  //    candidates (class > 0, a few percent of the pixels) are collected in a list for the hysteresis step
  for (int k = tid; k < TH * TW; k += 256) {
    int i = k / TW, j = k % TW;
    int r = r0 + i, c = c0 + j;
    s_par[k] = k;
    s_strong[k] = 0;
    if (r >= h || c >= w) {
      s_cls[i][j] = 0;
      continue;
    }
    uint8_t out = 0;
    if (r >= 1 && r < h - 1 && c >= 1 && c < w - 1) {
      float m = s_mag[i + 1][j + 1];
      if (!(m < t_low)) {
        const int dir = s_dir[i][j];
        float a = 0.f, b = 0.f;
        bool ok = true;
        switch (dir) {
          case 0: a = s_mag[i + 1][j]; b = s_mag[i + 1][j + 2]; break;      // (col-1,row) (col+1,row)
          case 45: a = s_mag[i][j]; b = s_mag[i + 2][j + 2]; break;         // (col-1,row-1) (col+1,row+1)
          case 90: a = s_mag[i][j + 1]; b = s_mag[i + 2][j + 1]; break;     // (col,row-1) (col,row+1)
          case 135: a = s_mag[i][j + 2]; b = s_mag[i + 2][j]; break;        // (col+1,row-1) (col-1,row+1)
          default: ok = false; break;
        }
        if (ok && m >= a && m >= b) out = (m >= t_high) ? 2 : 1;
      }
    }
    s_cls[i][j] = out;
    cls[(size_t)seg * stride + (size_t)r * w + c] = out;
    if (out) s_list[atomicAdd(&s_ncand, 1)] = (unsigned short)k;
  }
  __syncthreads();
  // 5. hysteresis, tile-local part: 8-connected components of the candidates inside this tile are merged in shared
  //    memory; every candidate then points straight at its tile root (the smallest pixel index of its local component,
  //    the same "larger under smaller" rule as the global merge -- so the result does not depend on the order of the list),
  //    and k_uf_border only has to stitch tile borders.  One thread per (candidate, backward neighbour).
  const int nc = s_ncand;
  for (int q = tid; q < 4 * nc; q += 256) {
    const int k = s_list[q >> 2], nb = q & 3;
    const int i = k / TW, j = k % TW;
    int other = -1;
    if (nb == 0) {
      if (j > 0 && s_cls[i][j - 1]) other = k - 1;
    } else if (i > 0) {
      const int jj = j + nb - 2;  // nb 1, 2, 3 -> columns j-1, j, j+1 of the row above
      if (jj >= 0 && jj < TW && s_cls[i - 1][jj]) other = k - TW + nb - 2;
    }
    if (other >= 0) uf_union(s_par, k, other);
  }
  __syncthreads();
  // every candidate records its tile root; a tile root is marked in its class byte (CLS_ROOT) together with "my local
  // component holds a strong pixel" (CLS_LSTRONG).  From here on only tile roots take part in the union-find: a non-root
  // candidate's parent is written once, here, and never touched again.
  int roots[4];
#pragma unroll
  for (int m = 0; m < 4; ++m) {
    const int q = tid + m * 256;
    roots[m] = -1;
    if (q < nc) {
      const int k = s_list[q];
      roots[m] = uf_find(s_par, k);
      if (s_cls[k / TW][k % TW] == 2) s_strong[roots[m]] = 1;
    }
  }
  __syncthreads();
#pragma unroll
  for (int m = 0; m < 4; ++m) {
    const int q = tid + m * 256;
    if (q < nc) {
      const int k = s_list[q], root = roots[m];
      const size_t gk = (size_t)seg * stride + (size_t)(r0 + k / TW) * w + (c0 + k % TW);
      parent[gk] = (r0 + root / TW) * w + (c0 + root % TW);
      if (root == k) cls[gk] = (uint8_t)(s_cls[k / TW][k % TW] | CLS_ROOT | (s_strong[k] ? CLS_LSTRONG : 0));
    }
  }
}

// the tile root of candidate p (class byte cb): itself, or the parent k_canny_nms wrote (never modified afterwards)
__device__ __forceinline__ int tile_root(const int* par, int p, unsigned cb) { return (cb & CLS_ROOT) ? p : par[p]; }

// stitch the tile-local components across tile borders (the only unions left): candidates in the first row, first column
// or last column of a k_canny_nms tile are united with their backward neighbours in global memory -- through their TILE
// ROOTS, so every path of the global forest runs over tile roots only
__global__ void k_uf_border(const uint8_t* __restrict__ cls, int* __restrict__ parent, int w, int h, int stride) {
  const int seg = blockIdx.y;
  const uint8_t* c = cls + (size_t)seg * stride;
  int* par = parent + (size_t)seg * stride;
  // border pixels only: per tile row TW pixels of the top row + 2 (TH - 1) of the side columns
  const int tiles_x = (w + TW - 1) / TW, tiles_y = (h + TH - 1) / TH;
  constexpr int PER = TW + 2 * (TH - 1);
  const long long total = (long long)tiles_x * tiles_y * PER;
  for (long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x; q < total; q += (long long)gridDim.x * blockDim.x) {
    const int t = (int)(q / PER), e = (int)(q % PER);
    const int ty = t / tiles_x, tx = t % tiles_x;
    int i, j;
    if (e < TW) {
      i = 0;
      j = e;
    } else {
      const int f = e - TW;
      i = 1 + (f >> 1);
      j = (f & 1) ? TW - 1 : 0;
    }
    const int r = ty * TH + i, col = tx * TW + j;
    if (r >= h || col >= w) continue;
    const int p = r * w + col;
    const unsigned cp = c[p];
    if (!cp) continue;
    // candidates are interior pixels, so the four backward neighbours are always in the image; neighbours of the same
    // tile share the tile root (already merged locally)
    const int rp = tile_root(par, p, cp);
    const int nbr[4] = {p - 1, p - w - 1, p - w, p - w + 1};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const unsigned cq = c[nbr[k]];
      if (!cq) continue;
      const int rq = tile_root(par, nbr[k], cq);
      if (rq != rp) uf_union(par, rp, rq);
    }
  }
}

// tile roots whose local component holds a strong pixel mark the root of their whole component
__global__ void k_uf_flag(const uint8_t* __restrict__ cls, int* __restrict__ parent, uint8_t* __restrict__ strong, int n,
                          int stride) {
  const int seg = blockIdx.y;
  const uint8_t* c = cls + (size_t)seg * stride;
  int* par = parent + (size_t)seg * stride;
  constexpr unsigned want = CLS_ROOT | CLS_LSTRONG;
  if ((stride & 15) == 0) {  // sixteen class bytes per load: tile roots are ~0.3 % of the pixels
    const uint4* c16 = reinterpret_cast<const uint4*>(c);
    for (int q = blockIdx.x * blockDim.x + threadIdx.x; 16 * q < n; q += gridDim.x * blockDim.x) {
      const uint4 v = c16[q];
      const unsigned wd[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (!(wd[u] & (wd[u] >> 1) & 0x04040404u)) continue;  // no byte with both bits
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int i = 16 * q + 4 * u + k;
          if ((((wd[u] >> (8 * k)) & want) == want) && i < n) strong[(size_t)seg * stride + uf_find(par, i)] = 1;
        }
      }
    }
    return;
  }
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    if ((c[i] & want) == want) strong[(size_t)seg * stride + uf_find(par, i)] = 1;
}

// ... and every tile root learns whether its whole component is strong (CLS_GSTRONG in its own class byte), so that
// k_edge_mask needs two dependent loads per candidate and no pointer chase
__global__ void k_root_pull(uint8_t* __restrict__ cls, int* __restrict__ parent, const uint8_t* __restrict__ strong, int n,
                            int stride) {
  const int seg = blockIdx.y;
  uint8_t* c = cls + (size_t)seg * stride;
  int* par = parent + (size_t)seg * stride;
  if ((stride & 15) == 0) {
    const uint4* c16 = reinterpret_cast<const uint4*>(c);
    for (int q = blockIdx.x * blockDim.x + threadIdx.x; 16 * q < n; q += gridDim.x * blockDim.x) {
      const uint4 v = c16[q];
      const unsigned wd[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (!(wd[u] & 0x08080808u)) continue;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int i = 16 * q + 4 * u + k;
          const unsigned b = (wd[u] >> (8 * k)) & 255u;
          if ((b & CLS_ROOT) && i < n && strong[(size_t)seg * stride + uf_find(par, i)]) c[i] = (uint8_t)(b | CLS_GSTRONG);
        }
      }
    }
    return;
  }
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const unsigned b = c[i];
    if ((b & CLS_ROOT) && strong[(size_t)seg * stride + uf_find(par, i)]) c[i] = (uint8_t)(b | CLS_GSTRONG);
  }
}

// pcl::OrganizedEdgeBase::extractEdges: NaN-boundary / occluding / occluded labels from the depth channel (the classes
// edge_extractor.hpp:26-34 gathers next to the RGB edges; one thread per interior pixel, the 3x3 neighbourhood is L1/L2
// traffic).  label |= 1 NAN_BOUNDARY, 2 OCCLUDING, 4 OCCLUDED.
__global__ void k_depth_edges(const float4* __restrict__ pts, int w, int h, int stride, float th, int max_search,
                              uint8_t* __restrict__ labels) {
  const int seg = blockIdx.y;
  const float4* P = pts + (size_t)seg * stride;
  uint8_t* L = labels + (size_t)seg * stride;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < w * h; idx += gridDim.x * blockDim.x) {
    const int row = idx / w, col = idx - row * w;
    uint8_t lab = 0;
    if (row >= 1 && row < h - 1 && col >= 1 && col < w - 1) {
      const float z = P[idx].z;
      if (isfinite(z)) {
        const float cd = fabsf(z);
        const int DX[8] = {-1, -1, 0, 1, 1, 1, 0, -1}, DY[8] = {0, -1, -1, -1, 0, 1, 1, 1};
        float nz[8];
        int dx = 0, dy = 0, n_inv = 0;
#pragma unroll
        for (int d = 0; d < 8; ++d) {
          nz[d] = P[idx + DY[d] * w + DX[d]].z;
          if (!isfinite(nz[d])) {
            dx += DX[d];
            dy += DY[d];
            ++n_inv;
          }
        }
        float mn = __fsub_rn(cd, fabsf(nz[0])), mx = mn;  // (only used when every neighbour is finite)
#pragma unroll
        for (int d = 1; d < 8; ++d) {
          const float dd = __fsub_rn(cd, fabsf(nz[d]));
          mn = dd < mn ? dd : mn;
          mx = dd > mx ? dd : mx;
        }
        if (n_inv == 0) {
          const float dom = fabsf(mn) > fabsf(mx) ? mn : mx;
          if (fabsf(dom) > fmul(th, cd)) lab |= dom > 0.f ? 4 : 2;
        } else {
          const float fdx = __fdiv_rn((float)dx, (float)n_inv), fdy = __fdiv_rn((float)dy, (float)n_inv);
          float corr = NAN;
          for (int s = 1; s < max_search; ++s) {
            const int sr = row + (int)floorf(fmul(fdy, (float)s)), sc = col + (int)floorf(fmul(fdx, (float)s));
            if (sr < 0 || sr >= h || sc < 0 || sc >= w) break;
            const float sz = P[sr * w + sc].z;
            if (isfinite(sz)) {
              corr = fabsf(sz);
              break;
            }
          }
          if (!isnan(corr)) {
            const float dd = __fsub_rn(cd, corr);
            if (fabsf(dd) > fmul(th, cd)) lab |= dd > 0.f ? 4 : 2;
          } else {
            lab |= 1;
          }
        }
      }
    }
    L[idx] = lab;
  }
}

__global__ void k_or_mask(const uint8_t* __restrict__ mask, uint8_t* __restrict__ labels, int n, int stride) {
  const int seg = blockIdx.y;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    if (mask[(size_t)seg * stride + i]) labels[(size_t)seg * stride + i] |= 16;  // EDGELABEL_RGB_CANNY
}

constexpr int CT = 256;       // threads per compaction block
constexpr int CPT = 16;       // consecutive pixels per thread (one 128-bit word of class / mask bytes)
constexpr int CB = CPT * CT;  // pixels per compaction block

// mask = candidate whose component holds a strong pixel; per-block edge counts for the ordered compaction.  Sixteen
// consecutive pixels per thread: ~97 % of the class words are zero, and these kernels are bound by the bytes (and the
// dependent loads of the few candidates) in flight, not by bandwidth.
__global__ void __launch_bounds__(CT) k_edge_mask(const uint8_t* __restrict__ cls, const int* __restrict__ parent,
                                                  uint8_t* __restrict__ mask, int* __restrict__ blk_cnt, int n, int stride,
                                                  int nblk) {
  __shared__ int s_w[CT / 32];
  const int seg = blockIdx.y;
  const int i0 = blockIdx.x * CB + CPT * threadIdx.x;
  int cnt = 0;
  const uint8_t* c = cls + (size_t)seg * stride;
  const int* par = parent + (size_t)seg * stride;
  // a candidate is an edge iff the component of its tile root is strong: its own byte if it is the root, else the root's
  auto is_edge = [&](int i, unsigned b) -> bool {
    if (!(b & 3u)) return false;
    if (b & CLS_ROOT) return (b & CLS_GSTRONG) != 0;
    return (c[par[i]] & CLS_GSTRONG) != 0;
  };
  if (i0 < n) {
    const size_t g0 = (size_t)seg * stride + i0;
    if ((stride & 15) == 0 && i0 + CPT - 1 < n) {
      const uint4 v = *reinterpret_cast<const uint4*>(cls + g0);
      const unsigned wd[4] = {v.x, v.y, v.z, v.w};
      unsigned out[4] = {0u, 0u, 0u, 0u};
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (!(wd[u] & 0x03030303u)) continue;
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if (is_edge(i0 + 4 * u + k, (wd[u] >> (8 * k)) & 255u)) {
            out[u] |= 255u << (8 * k);
            ++cnt;
          }
      }
      *reinterpret_cast<uint4*>(mask + g0) = make_uint4(out[0], out[1], out[2], out[3]);
    } else {
      for (int k = 0; k < CPT && i0 + k < n; ++k) {
        const int e = is_edge(i0 + k, c[i0 + k]) ? 1 : 0;
        mask[g0 + k] = e ? 255 : 0;
        cnt += e;
      }
    }
  }
  cnt = __reduce_add_sync(0xffffffffu, cnt);
  if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = cnt;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
#pragma unroll
    for (int w = 0; w < CT / 32; ++w) t += s_w[w];
    blk_cnt[seg * nblk + blockIdx.x] = t;
  }
}

// one CTA per frame: exclusive scan of its block counts (nblk <= a few thousand), writes the frame's edge count
__global__ void __launch_bounds__(1024) k_seg_scan(int* __restrict__ blk_cnt, int nblk, int* __restrict__ out_count,
                                                   int out_stride, int* __restrict__ overflow) {
  __shared__ int warp_tot[32];
  __shared__ int carry_s;
  const int seg = blockIdx.x;
  int* b = blk_cnt + seg * nblk;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  for (int base = 0; base < nblk; base += 1024) {
    int i = base + threadIdx.x;
    int v = i < nblk ? b[i] : 0;
    int incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    if (lane == 31) warp_tot[wid] = incl;
    __syncthreads();
    if (wid == 0) {
      int wv = warp_tot[lane], wi = wv;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, wi, o);
        if (lane >= o) wi += t;
      }
      warp_tot[lane] = wi - wv;
    }
    __syncthreads();
    int carry = carry_s;
    int excl = carry + warp_tot[wid] + incl - v;
    if (i < nblk) b[i] = excl;
    __syncthreads();
    if (threadIdx.x == 1023) carry_s = excl + v;
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    int total = carry_s;
    if (total > out_stride) {
      atomicExch(overflow, 1);
      total = 0;
    }
    out_count[seg] = total;
  }
}

__global__ void __launch_bounds__(CT) k_scatter(const uint8_t* __restrict__ mask, const int* __restrict__ blk_off,
                                                const float4* __restrict__ pts, const int* __restrict__ out_count,
                                                float4* __restrict__ out, int n, int stride, int out_stride, int nblk) {
  __shared__ int warp_tot[CT / 32];
  const int seg = blockIdx.y;
  if (out_count[seg] == 0) return;  // empty or overflowed frame
  const int i0 = blockIdx.x * CB + CPT * threadIdx.x;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  unsigned flags = 0u;  // bit k: pixel i0 + k is an edge
  if (i0 < n) {
    const size_t g0 = (size_t)seg * stride + i0;
    if ((stride & 15) == 0 && i0 + CPT - 1 < n) {
      const uint4 v = *reinterpret_cast<const uint4*>(mask + g0);
      const unsigned wd[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int k = 0; k < 4; ++k) flags |= ((wd[u] >> (8 * k)) & 255u) ? (1u << (4 * u + k)) : 0u;
    } else {
      for (int k = 0; k < CPT && i0 + k < n; ++k) flags |= mask[g0 + k] ? (1u << k) : 0u;
    }
  }
  const int c = __popc(flags);
  int incl = c;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) warp_tot[wid] = incl;
  __syncthreads();
  if (flags) {  // ascending pixel order: block offset + warps before + lanes before + own earlier pixels
    int pos = blk_off[seg * nblk + blockIdx.x] + incl - c;
#pragma unroll
    for (int w = 0; w < CT / 32; ++w) pos += (w < wid) ? warp_tot[w] : 0;
    while (flags) {
      const int k = __ffs(flags) - 1;
      flags &= flags - 1;
      out[(size_t)seg * out_stride + pos++] = pts[(size_t)seg * stride + i0 + k];
    }
  }
}

}  // namespace

static int edge_extract_impl(rspcl_ctx* ctx, const rspcl_cloud* frames, float t_low, float t_high, rspcl_cloud* out_edges,
                             uint8_t* host_mask, uint8_t* d_labels_or);

extern "C" int rspcl_edge_extract(rspcl_ctx* ctx, const rspcl_cloud* frames, float t_low, float t_high,
                                  rspcl_cloud* out_edges, uint8_t* host_mask) {
  return edge_extract_impl(ctx, frames, t_low, t_high, out_edges, host_mask, nullptr);
}

// pcl::OrganizedEdgeFromRGBNormals::compute as edge_extractor.hpp:17-24 configures it, minus the HIGH_CURVATURE class (which
// needs the integral-image normals): per pixel label = 1 NAN_BOUNDARY | 2 OCCLUDING | 4 OCCLUDED | 16 RGB_CANNY.
extern "C" int rspcl_edge_labels(rspcl_ctx* ctx, const rspcl_cloud* frames, float th_depth_discon, int max_search_neighbors,
                                 float t_low, float t_high, uint8_t* host_labels) {
  if (!ctx || !frames || !host_labels) return RSPCL_ERR_ARG;
  if (frames->height <= 0 || frames->width <= 0) RSPCL_FAIL(ctx, RSPCL_ERR_ARG, "edge_labels: input is not organized");
  CU(ctx, cudaSetDevice(ctx->device));
  const int w = frames->width, h = frames->height, n = w * h, S = frames->n_seg;
  Scratch scr(ctx);
  uint8_t* d_lab = nullptr;
  CU(ctx, scr.alloc(&d_lab, (size_t)S * frames->stride));
  dim3 g(blocks_per_seg(ctx, S, n, 256), S);
  k_depth_edges<<<g, 256, 0, ctx->stream>>>(frames->pts, w, h, frames->stride, th_depth_discon, max_search_neighbors, d_lab);
  LAUNCH_CHECK(ctx);
  TmpCloud E(ctx);
  int rc = E.init(S, n);
  if (rc) RSPCL_FAIL(ctx, rc, "edge_labels: scratch allocation failed");
  rc = edge_extract_impl(ctx, frames, t_low, t_high, &E.c, nullptr, d_lab);
  if (rc) return rc;
  for (int s = 0; s < S; ++s) CU(ctx, small_d2h(ctx, host_labels + (size_t)s * n, d_lab + (size_t)s * frames->stride, n));
  CU(ctx, ctx_sync(ctx));
  scr.ok();
  return RSPCL_OK;
}

static int edge_extract_impl(rspcl_ctx* ctx, const rspcl_cloud* frames, float t_low, float t_high, rspcl_cloud* out_edges,
                             uint8_t* host_mask, uint8_t* d_labels_or) {
  if (!ctx || !frames || !out_edges || frames == out_edges) return RSPCL_ERR_ARG;
  if (frames->height <= 0 || frames->width <= 0) RSPCL_FAIL(ctx, RSPCL_ERR_ARG, "edge_extract: input is not organized");
  if (out_edges->n_seg != frames->n_seg) RSPCL_FAIL(ctx, RSPCL_ERR_ARG, "edge_extract: output n_seg mismatch");
  CU(ctx, cudaSetDevice(ctx->device));
  const int w = frames->width, h = frames->height, n = w * h, S = frames->n_seg, stride = frames->stride;
  // a stale or missing gray plane is not rebuilt: k_canny_nms<true> reads the colours of the points directly
  const bool from_pts = !(frames->gray && frames->gray_valid);
  const size_t tot = (size_t)S * stride;
  const int nblk = div_up(n, CB);
  Scratch scr(ctx);  // released on every exit path
  uint8_t *cls = nullptr, *strong = nullptr, *mask = nullptr;
  int *parent = nullptr, *blk = nullptr, *d_over = nullptr;
  CU(ctx, scr.alloc(&cls, tot));
  CU(ctx, scr.alloc(&strong, tot));
  CU(ctx, scr.alloc(&mask, tot));
  CU(ctx, scr.alloc(&parent, tot));
  CU(ctx, scr.alloc(&blk, (size_t)S * nblk));
  CU(ctx, scr.alloc(&d_over, 1));
  CU(ctx, cudaMemsetAsync(strong, 0, tot, ctx->stream));
  CU(ctx, cudaMemsetAsync(d_over, 0, sizeof(int), ctx->stream));

  dim3 g1(div_up(w, TW), div_up(h, TH), S);
  {
    ProfScope prof(ctx, "k_canny_nms", (double)S * n);
    if (from_pts)
      k_canny_nms<true><<<g1, dim3(64, 4), 0, ctx->stream>>>(nullptr, frames->pts, w, h, stride, t_low, t_high, cls, parent);
    else
      k_canny_nms<false><<<g1, dim3(64, 4), 0, ctx->stream>>>(frames->gray, nullptr, w, h, stride, t_low, t_high, cls, parent);
    LAUNCH_CHECK(ctx);
  }
  dim3 g2(blocks_per_seg(ctx, S, n, 256), S);
  ProfScope prof_h(ctx, "edge_hysteresis_compact", (double)S * n);
  // latency-bound scans (one dependent chain per hit): one item per thread, no grid-stride loop, so that as many chains as
  // possible are in flight
  const long long border_items = (long long)div_up(w, TW) * div_up(h, TH) * (TW + 2 * (TH - 1));
  dim3 gb((unsigned)div_up(border_items, 256), S), gf((unsigned)div_up(div_up(n, 16), 256), S);
  k_uf_border<<<gb, 256, 0, ctx->stream>>>(cls, parent, w, h, stride);
  LAUNCH_CHECK(ctx);
  k_uf_flag<<<gf, 256, 0, ctx->stream>>>(cls, parent, strong, n, stride);
  LAUNCH_CHECK(ctx);
  k_root_pull<<<gf, 256, 0, ctx->stream>>>(cls, parent, strong, n, stride);
  LAUNCH_CHECK(ctx);
  dim3 g3(nblk, S);
  k_edge_mask<<<g3, CT, 0, ctx->stream>>>(cls, parent, mask, blk, n, stride, nblk);
  LAUNCH_CHECK(ctx);
  k_seg_scan<<<S, 1024, 0, ctx->stream>>>(blk, nblk, out_edges->count, out_edges->stride, d_over);
  LAUNCH_CHECK(ctx);
  k_scatter<<<g3, CT, 0, ctx->stream>>>(mask, blk, frames->pts, out_edges->count, out_edges->pts, n, stride,
                                        out_edges->stride, nblk);
  LAUNCH_CHECK(ctx);
  out_edges->width = out_edges->height = 0;
  out_edges->max_count_hint = out_edges->stride < n ? out_edges->stride : n;
  if (d_labels_or) {
    k_or_mask<<<g2, 256, 0, ctx->stream>>>(mask, d_labels_or, n, stride);
    LAUNCH_CHECK(ctx);
  }

  int over = 0;
  if (host_mask) {
    for (int s = 0; s < S; ++s)
      CU(ctx, small_d2h(ctx, host_mask + (size_t)s * n, mask + (size_t)s * stride, n));
    CU(ctx, small_d2h(ctx, &over, d_over, sizeof(int)));
    CU(ctx, ctx_sync(ctx));
  } else if (out_edges->stride < n) {
    // capacity below the worst case: the overflow flag must be checked (synchronises)
    CU(ctx, small_d2h(ctx, &over, d_over, sizeof(int)));
    CU(ctx, ctx_sync(ctx));
  }
  if (over) RSPCL_FAIL(ctx, RSPCL_ERR_CAPACITY, "edge_extract: a frame produced more edge points than the output stride %d", out_edges->stride);
  scr.ok();
  return RSPCL_OK;
}

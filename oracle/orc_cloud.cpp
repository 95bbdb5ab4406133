// orc_cloud.cpp -- CPU oracle (test infrastructure only): ApproximateVoxelGrid, transformPointCloud,
// exact nearest neighbour (brute force and kd-tree).
//
// Reference call sites: icp:37,47,59-60,75-76; ndt:34,45,57-58,68-69; incr:36,54-55 (voxel filter);
// icp:116-117, ndt:104-105, incr:63 (transformPointCloud).  PCL 1.9.1 sources followed (not vendored --
// PARITY UNPINNED, see orc.h): filters/impl/approximate_voxel_grid.hpp, common/impl/transforms.hpp,
// kdtree/impl/kdtree_flann.hpp + FLANN L2_Simple<float>.
#include "orc.h"
#include <cmath>
#include <climits>
#include <cstring>
#include <vector>
#include <algorithm>
#include <numeric>

namespace {

// x86-64 cvttss2si semantics for static_cast<int>(floor(v)): NaN / out of range -> INT_MIN ("integer indefinite")
inline int floor_to_int(float v) {
  float f = std::floor(v);
  if (!(f >= -2147483648.0f && f < 2147483648.0f)) return INT_MIN;
  return static_cast<int>(f);
}

struct HistEntry {
  int ix, iy, iz, count;
  float c[7];  // x, y, z, rgb-as-float (unused garbage in PCL), r, g, b
};

}  // namespace

extern "C" void orc_voxel_keys(const OrcPoint* in, int n, const float leaf[3], int32_t* ijk, int32_t* slot) {
  const float inv[3] = {1.0f / leaf[0], 1.0f / leaf[1], 1.0f / leaf[2]};
  for (int i = 0; i < n; ++i) {
    int ix = floor_to_int(in[i].x * inv[0]);
    int iy = floor_to_int(in[i].y * inv[1]);
    int iz = floor_to_int(in[i].z * inv[2]);
    unsigned hash = (unsigned(ix) * 7171u + unsigned(iy) * 3079u + unsigned(iz) * 4231u) & 511u;
    if (ijk) {
      ijk[3 * i] = ix;
      ijk[3 * i + 1] = iy;
      ijk[3 * i + 2] = iz;
    }
    if (slot) slot[i] = int(hash);
  }
}

extern "C" int orc_approx_voxel(const OrcPoint* in, int n, const float leaf[3], OrcPoint* out) {
  // approximate_voxel_grid.hpp applyFilter: 512-entry direct-mapped history; a colliding different voxel
  // flushes the slot (centroid = sum / float(count)); remaining slots are flushed in slot order.
  const int histsize = 512;
  std::vector<HistEntry> hist(histsize);
  for (auto& e : hist) {
    e.count = 0;
    for (float& v : e.c) v = 0.f;
  }
  const float inv[3] = {1.0f / leaf[0], 1.0f / leaf[1], 1.0f / leaf[2]};
  int op = 0;
  auto flush = [&](HistEntry& e) {
    float cnt = static_cast<float>(e.count);
    float x = e.c[0] / cnt, y = e.c[1] / cnt, z = e.c[2] / cnt;
    float r = e.c[4] / cnt, g = e.c[5] / cnt, b = e.c[6] / cnt;
    uint32_t rgb = (uint32_t(int(r)) << 16) | (uint32_t(int(g)) << 8) | uint32_t(int(b));  // alpha byte 0
    out[op++] = OrcPoint{x, y, z, rgb};
  };
  for (int cp = 0; cp < n; ++cp) {
    int ix = floor_to_int(in[cp].x * inv[0]);
    int iy = floor_to_int(in[cp].y * inv[1]);
    int iz = floor_to_int(in[cp].z * inv[2]);
    unsigned hash = (unsigned(ix) * 7171u + unsigned(iy) * 3079u + unsigned(iz) * 4231u) & (histsize - 1);
    HistEntry& e = hist[hash];
    if (e.count && (ix != e.ix || iy != e.iy || iz != e.iz)) {
      flush(e);
      e.count = 0;
      for (float& v : e.c) v = 0.f;
    }
    e.ix = ix;
    e.iy = iy;
    e.iz = iz;
    e.count++;
    uint32_t c = in[cp].rgba;
    e.c[0] += in[cp].x;
    e.c[1] += in[cp].y;
    e.c[2] += in[cp].z;
    e.c[4] += float((c >> 16) & 255);
    e.c[5] += float((c >> 8) & 255);
    e.c[6] += float(c & 255);
  }
  for (int i = 0; i < histsize; ++i)
    if (hist[i].count) flush(hist[i]);
  return op;
}

extern "C" void orc_transform(const OrcPoint* in, int n, const float T[16], OrcPoint* out) {
  // transforms.hpp (Matrix4 overload): x' = m00*x + m01*y + m02*z + m03, left-to-right in float; rgb copied.
  // Column-major T: m(r,c) = T[c*4 + r].
  for (int i = 0; i < n; ++i) {
    float x = in[i].x, y = in[i].y, z = in[i].z;
    OrcPoint o;
    o.rgba = in[i].rgba;
    if (!std::isfinite(x) || !std::isfinite(y) || !std::isfinite(z)) {  // is_dense=false branch: copied, not moved
      o.x = x;
      o.y = y;
      o.z = z;
    } else {
      o.x = T[0] * x + T[4] * y + T[8] * z + T[12];
      o.y = T[1] * x + T[5] * y + T[9] * z + T[13];
      o.z = T[2] * x + T[6] * y + T[10] * z + T[14];
    }
    out[i] = o;
  }
}

// FLANN L2_Simple<float>: result += diff*diff over x,y,z sequentially
static inline float dist2(const OrcPoint& a, const OrcPoint& b) {
  float r = 0.f, d;
  d = a.x - b.x;
  r += d * d;
  d = a.y - b.y;
  r += d * d;
  d = a.z - b.z;
  r += d * d;
  return r;
}

static inline bool finite_pt(const OrcPoint& p) { return std::isfinite(p.x) && std::isfinite(p.y) && std::isfinite(p.z); }

extern "C" void orc_nn_brute(const OrcPoint* tgt, int nt, const OrcPoint* q, int nq, int32_t* idx, float* d2) {
  for (int i = 0; i < nq; ++i) {
    int best = -1;
    float bd = INFINITY;
    if (finite_pt(q[i]))
      for (int j = 0; j < nt; ++j) {
        if (!finite_pt(tgt[j])) continue;
        float d = dist2(q[i], tgt[j]);
        if (d < bd) {  // strict: lowest index wins ties
          bd = d;
          best = j;
        }
      }
    idx[i] = best;
    d2[i] = bd;
  }
}

namespace {
// kd-tree over finite target points; leaves of <= 15 points like PCL's KDTreeSingleIndexParams(15).
// Exactness: a far subtree is skipped only if fl(diff*diff) > best, and fl(dx*dx) <= full float d2 because
// float addition of non-negative terms is monotone, so ties are always explored -> lowest index wins.
struct KdTree {
  struct Node {
    int lo, hi;       // point range [lo,hi) in perm
    int axis;         // -1 for leaf
    float split;
    int left, right;
  };
  const OrcPoint* pts;
  std::vector<int> perm;
  std::vector<Node> nodes;
  std::vector<float> xs, ys, zs;  // leaf-ordered coordinates
  int build(int lo, int hi) {
    int id = int(nodes.size());
    nodes.push_back(Node{lo, hi, -1, 0.f, -1, -1});
    if (hi - lo <= 15) return id;
    float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (int i = lo; i < hi; ++i) {
      const OrcPoint& p = pts[perm[i]];
      float v[3] = {p.x, p.y, p.z};
      for (int a = 0; a < 3; ++a) {
        mn[a] = std::min(mn[a], v[a]);
        mx[a] = std::max(mx[a], v[a]);
      }
    }
    int axis = 0;
    if (mx[1] - mn[1] > mx[axis] - mn[axis]) axis = 1;
    if (mx[2] - mn[2] > mx[axis] - mn[axis]) axis = 2;
    if (!(mx[axis] > mn[axis])) return id;  // all identical: keep as (big) leaf
    int mid = (lo + hi) / 2;
    auto key = [&](int i) {
      const OrcPoint& p = pts[i];
      return axis == 0 ? p.x : (axis == 1 ? p.y : p.z);
    };
    std::nth_element(perm.begin() + lo, perm.begin() + mid, perm.begin() + hi,
                     [&](int a, int b) { return key(a) < key(b); });
    float split = key(perm[mid]);
    int l = build(lo, mid);
    int r = build(mid, hi);
    nodes[id].axis = axis;
    nodes[id].split = split;
    nodes[id].left = l;
    nodes[id].right = r;
    return id;
  }
  void init(const OrcPoint* p, int n) {
    pts = p;
    perm.clear();
    for (int i = 0; i < n; ++i)
      if (finite_pt(p[i])) perm.push_back(i);
    nodes.clear();
    nodes.reserve(perm.size() / 4 + 8);
    if (!perm.empty()) build(0, int(perm.size()));
    xs.resize(perm.size());
    ys.resize(perm.size());
    zs.resize(perm.size());
    for (size_t i = 0; i < perm.size(); ++i) {
      xs[i] = p[perm[i]].x;
      ys[i] = p[perm[i]].y;
      zs[i] = p[perm[i]].z;
    }
  }
  void search(int nid, const OrcPoint& q, int& best, float& bd) const {
    const Node& nd = nodes[nid];
    if (nd.axis < 0) {
      for (int i = nd.lo; i < nd.hi; ++i) {
        float r = 0.f, d;
        d = q.x - xs[i];
        r += d * d;
        d = q.y - ys[i];
        r += d * d;
        d = q.z - zs[i];
        r += d * d;
        int j = perm[i];
        if (r < bd || (r == bd && j < best)) {
          bd = r;
          best = j;
        }
      }
      return;
    }
    float qv = nd.axis == 0 ? q.x : (nd.axis == 1 ? q.y : q.z);
    float diff = qv - nd.split;
    int first = diff < 0 ? nd.left : nd.right;
    int second = diff < 0 ? nd.right : nd.left;
    search(first, q, best, bd);
    if (!(diff * diff > bd)) search(second, q, best, bd);
  }
  void query(const OrcPoint& q, int& best, float& bd) const {
    best = -1;
    bd = INFINITY;
    if (nodes.empty() || !finite_pt(q)) return;
    search(0, q, best, bd);
  }
};
}  // namespace

// Opaque handle used by the ICP / fitness code in orc_icp.cpp
struct OrcKd {
  KdTree t;
};
OrcKd* orc_kd_build(const OrcPoint* tgt, int nt) {
  OrcKd* k = new OrcKd;
  k->t.init(tgt, nt);
  return k;
}
void orc_kd_free(OrcKd* k) { delete k; }
void orc_kd_query(const OrcKd* k, const OrcPoint* q, int* idx, float* d2) { k->t.query(*q, *idx, *d2); }

extern "C" void orc_nn_kdtree(const OrcPoint* tgt, int nt, const OrcPoint* q, int nq, int32_t* idx, float* d2) {
  KdTree t;
  t.init(tgt, nt);
  for (int i = 0; i < nq; ++i) {
    int b;
    float d;
    t.query(q[i], b, d);
    idx[i] = b;
    d2[i] = d;
  }
}

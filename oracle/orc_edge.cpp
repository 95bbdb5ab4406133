// orc_edge.cpp -- CPU oracle (test infrastructure only): RGB-Canny edge extraction and the 3/5 crop.
//
// Follows, on the reference side, src/edge_extractor.hpp:7-39 (only label_indices[4], the RGB-Canny
// class, reaches the caller: edge_extractor.hpp:36-38) and src/blur_filter.hpp:18-36; on the PCL side
// (1.9.1, not vendored -- PARITY UNPINNED, see orc.h):
//   features/impl/organized_edge_detection.hpp  OrganizedEdgeFromRGB::extractEdges
//   2d/impl/edge.hpp        Edge::detectEdgeCanny / detectEdgeSobel / discretizeAngles /
//                           suppressNonMaxima / cannyTraceEdge
//   2d/impl/kernel.hpp      gaussianKernel / sobelKernelX / sobelKernelY
//   2d/impl/convolution.hpp Convolution::filter, BOUNDARY_OPTION_CLAMP
// Compile with -ffp-contract=off: PCL's header templates run as plain IEEE mul/add on x86-64.
#include "orc.h"
#include <cmath>
#include <limits>
#include <cfloat>
#include <cstring>
#include <vector>

extern "C" void orc_gaussian_kernel3(float k[9]) {
  // kernel.hpp gaussianKernel: kernel(j,i) = expf(-(iks^2+jks^2)/(2 sigma^2)), float running sum, divide.
  const int ks = 3;
  const float sigma = 1.0f;
  float sum = 0;
  double sigma_sqr = 2 * sigma * sigma;
  for (int i = 0; i < ks; i++)
    for (int j = 0; j < ks; j++) {
      int iks = i - ks / 2, jks = j - ks / 2;
      k[i * ks + j] = expf(float(-double(iks * iks + jks * jks) / sigma_sqr));
      sum += k[i * ks + j];
    }
  for (int i = 0; i < 9; ++i) k[i] /= sum;
}

// convolution.hpp: out(j,i) = sum_k sum_l kernel(l,k) * in(clamp(j+l-1), clamp(i+k-1)); k (row) outer.
static void conv3_clamp(const float* in, int w, int h, const float* kern /* [row k][col l] */, float* out) {
  for (int i = 0; i < h; i++)
    for (int j = 0; j < w; j++) {
      float intensity = 0;
      for (int k = 0; k < 3; k++)
        for (int l = 0; l < 3; l++) {
          int r = i + k - 1, c = j + l - 1;
          r = r < 0 ? 0 : (r >= h ? h - 1 : r);
          c = c < 0 ? 0 : (c >= w ? w - 1 : c);
          intensity += kern[k * 3 + l] * in[r * w + c];
        }
      out[i * w + j] = intensity;
    }
}

static inline uint8_t discretize_angle(float direction_rad, int* near_edge) {
  // edge.hpp discretizeAngles with pcl::rad2deg(float) = alpha * 57.29578f
  float angle = direction_rad * 57.29578f;
  if (near_edge) {
    const float th[8] = {22.5f, 67.5f, 112.5f, 157.5f, -22.5f, -67.5f, -112.5f, -157.5f};
    for (int i = 0; i < 8; ++i)
      if (std::fabs(angle - th[i]) < 1e-3f) *near_edge = 1;
  }
  if (((angle <= 22.5) && (angle >= -22.5)) || (angle >= 157.5) || (angle <= -157.5)) return 0;
  if (((angle > 22.5) && (angle < 67.5)) || ((angle < -112.5) && (angle > -157.5))) return 45;
  if (((angle >= 67.5) && (angle <= 112.5)) || ((angle <= -67.5) && (angle >= -112.5))) return 90;
  if (((angle > 112.5) && (angle < 157.5)) || ((angle < -22.5) && (angle > -67.5))) return 135;
  return 255;  // NaN direction: PCL leaves the raw radian value, int() of which hits no switch case
}

extern "C" int orc_canny(const OrcPoint* cloud, int w, int h, float t_low, float t_high, uint8_t* mask,
                         float* dbg_blur, float* dbg_gx, float* dbg_gy, float* dbg_mag, uint8_t* dbg_dir,
                         float* dbg_maxima, int* near_bin_edge) {
  const size_t n = size_t(w) * h;
  std::vector<float> gray(n), blur(n), gx(n), gy(n), mag(n), maxima(n, 0.0f);
  std::vector<uint8_t> dir(n);
  // organized_edge_detection.hpp: gray = float((r + g + b) / 3) with integer division
  for (size_t i = 0; i < n; ++i) {
    uint32_t c = cloud[i].rgba;
    int r = (c >> 16) & 255, g = (c >> 8) & 255, b = c & 255;
    gray[i] = float((r + g + b) / 3);
  }
  float gk[9];
  orc_gaussian_kernel3(gk);
  conv3_clamp(gray.data(), w, h, gk, blur.data());
  // kernel.hpp sobelKernelX / sobelKernelY, indexed [row][col]
  const float sx[9] = {-1, 0, 1, -2, 0, 2, -1, 0, 1};
  const float sy[9] = {-1, -2, -1, 0, 0, 0, 1, 2, 1};
  conv3_clamp(blur.data(), w, h, sx, gx.data());
  conv3_clamp(blur.data(), w, h, sy, gy.data());
  int near_cnt = 0;
  for (size_t i = 0; i < n; ++i) {
    mag[i] = std::sqrt(gx[i] * gx[i] + gy[i] * gy[i]);
    int ne = 0;
    dir[i] = discretize_angle(atan2f(gy[i], gx[i]), &ne);
    if (ne && mag[i] >= t_low) near_cnt++;
  }
  // suppressNonMaxima: interior only, skip mag < tLow, keep if >= both neighbours along the direction
  for (int i = 1; i < h - 1; i++)
    for (int j = 1; j < w - 1; j++) {
      float m = mag[i * w + j];
      if (m < t_low) continue;
      auto M = [&](int col, int row) { return mag[row * w + col]; };
      switch (dir[i * w + j]) {
        case 0:
          if (m >= M(j - 1, i) && m >= M(j + 1, i)) maxima[i * w + j] = m;
          break;
        case 45:
          if (m >= M(j - 1, i - 1) && m >= M(j + 1, i + 1)) maxima[i * w + j] = m;
          break;
        case 90:
          if (m >= M(j, i - 1) && m >= M(j, i + 1)) maxima[i * w + j] = m;
          break;
        case 135:
          if (m >= M(j + 1, i - 1) && m >= M(j - 1, i + 1)) maxima[i * w + j] = m;
          break;
        default:
          break;
      }
    }
  if (dbg_maxima) memcpy(dbg_maxima, maxima.data(), n * sizeof(float));
  // hysteresis: cannyTraceEdge recursion restated with an explicit stack (same visited set: 8-neighbour
  // flood through pixels that are neither 0 nor marked, bounds row>0,col>0,row<h,col<w)
  std::vector<int> stack;
  for (int i = 0; i < h; i++)
    for (int j = 0; j < w; j++) {
      float v = maxima[i * w + j];
      if (v < t_high || v == FLT_MAX) continue;
      maxima[i * w + j] = FLT_MAX;
      stack.push_back(i * w + j);
      while (!stack.empty()) {
        int p = stack.back();
        stack.pop_back();
        int pr = p / w, pc = p % w;
        for (int dr = -1; dr <= 1; ++dr)
          for (int dc = -1; dc <= 1; ++dc) {
            if (!dr && !dc) continue;
            int nr = pr + dr, nc = pc + dc;
            if (nr > 0 && nr < h && nc > 0 && nc < w) {
              float& pt = maxima[nr * w + nc];
              if (pt == 0.0f || pt == FLT_MAX) continue;
              pt = FLT_MAX;
              stack.push_back(nr * w + nc);
            }
          }
      }
    }
  int cnt = 0;
  for (size_t i = 0; i < n; ++i) {
    uint8_t e = maxima[i] == FLT_MAX ? 255 : 0;
    if (mask) mask[i] = e;
    cnt += e != 0;
  }
  if (dbg_blur) memcpy(dbg_blur, blur.data(), n * sizeof(float));
  if (dbg_gx) memcpy(dbg_gx, gx.data(), n * sizeof(float));
  if (dbg_gy) memcpy(dbg_gy, gy.data(), n * sizeof(float));
  if (dbg_mag) memcpy(dbg_mag, mag.data(), n * sizeof(float));
  if (dbg_dir) memcpy(dbg_dir, dir.data(), n);
  if (near_bin_edge) *near_bin_edge = near_cnt;
  return cnt;
}

extern "C" int orc_extract_edges(const OrcPoint* cloud, int w, int h, float t_low, float t_high, OrcPoint* out,
                                 int32_t* out_idx) {
  // edge_extractor.hpp:36: copyPointCloud(cloud, label_indices[4]) -- ascending row-major pixel indices
  std::vector<uint8_t> mask(size_t(w) * h);
  orc_canny(cloud, w, h, t_low, t_high, mask.data(), 0, 0, 0, 0, 0, 0, 0);
  int n = 0;
  for (int i = 0; i < w * h; ++i)
    if (mask[i]) {
      if (out) out[n] = cloud[i];
      if (out_idx) out_idx[n] = i;
      ++n;
    }
  return n;
}

extern "C" void orc_depth_edge_labels(const OrcPoint* cloud, int w, int h, float th, int max_search, uint8_t* labels) {
  // OrganizedEdgeBase<PointT, PointLT>::extractEdges: interior pixels only; eight neighbours in the order
  // (-1,0) (-1,-1) (0,-1) (1,-1) (1,0) (1,1) (0,1) (-1,1) (dx, dy).  A finite pixel whose neighbours are all finite is
  // labelled by the dominant (largest |.|) depth difference curr - neighbour against th * curr; a finite pixel next
  // to non-finite ones searches across the hole in the mean direction of the invalid neighbours.
  static const int DX[8] = {-1, -1, 0, 1, 1, 1, 0, -1}, DY[8] = {0, -1, -1, -1, 0, 1, 1, 1};
  for (int i = 0; i < w * h; ++i) labels[i] = 0;
  for (int row = 1; row < h - 1; ++row)
    for (int col = 1; col < w - 1; ++col) {
      const int cur = row * w + col;
      if (!std::isfinite(cloud[cur].z)) continue;
      const float cd = std::fabs(cloud[cur].z);
      float dist[8];
      bool invalid = false;
      for (int d = 0; d < 8; ++d) {
        const float nz = cloud[cur + DY[d] * w + DX[d]].z;
        if (!std::isfinite(nz)) {
          invalid = true;
          break;
        }
        dist[d] = cd - std::fabs(nz);
      }
      if (!invalid) {
        float mn = dist[0], mx = dist[0];
        for (int d = 1; d < 8; ++d) {
          mn = dist[d] < mn ? dist[d] : mn;
          mx = dist[d] > mx ? dist[d] : mx;
        }
        const float dom = std::fabs(mn) > std::fabs(mx) ? mn : mx;
        if (std::fabs(dom) > th * std::fabs(cd)) labels[cur] |= dom > 0.f ? 4 : 2;  // OCCLUDED : OCCLUDING
      } else {
        int dx = 0, dy = 0, n_inv = 0;
        for (int d = 0; d < 8; ++d)
          if (!std::isfinite(cloud[cur + DY[d] * w + DX[d]].z)) {
            dx += DX[d];
            dy += DY[d];
            ++n_inv;
          }
        const float fdx = float(dx) / float(n_inv), fdy = float(dy) / float(n_inv);
        float corr = std::numeric_limits<float>::quiet_NaN();
        for (int s = 1; s < max_search; ++s) {
          const int sr = row + int(std::floor(fdy * float(s))), sc = col + int(std::floor(fdx * float(s)));
          if (sr < 0 || sr >= h || sc < 0 || sc >= w) break;
          if (std::isfinite(cloud[sr * w + sc].z)) {
            corr = std::fabs(cloud[sr * w + sc].z);
            break;
          }
        }
        if (!std::isnan(corr)) {
          const float dd = cd - corr;
          if (std::fabs(dd) > th * std::fabs(cd)) labels[cur] |= dd > 0.f ? 4 : 2;
        } else {
          labels[cur] |= 1;  // NAN_BOUNDARY
        }
      }
    }
}

extern "C" int orc_crop35(const OrcPoint* cloud, int w, int h, OrcPoint* out, int* out_w, int* out_h) {
  // blur_filter.hpp:23-35.  NOTE the loop bounds (h/5 .. h/5*4, w/5 .. w/5*4) can produce fewer points than
  // (w*3/5)*(h*3/5) when w or h is not a multiple of 5 (e.g. RealSense 848x480: 507 copied columns, width 508).
  // input_cloud->points.resize() SHRINKS the vector (blur_filter.hpp:25), so the tail [n_copied, ow*oh) keeps the
  // points the input held at those indices before the call.
  int ow = w * 3 / 5, oh = h * 3 / 5;
  for (int k = 0; k < ow * oh; ++k) out[k] = cloud[k];
  int i = 0;
  for (int r = h / 5; r < h / 5 * 4; r++)
    for (int c = w / 5; c < w / 5 * 4; c++) {
      if (i < ow * oh) out[i] = cloud[r * w + c];
      i++;
    }
  *out_w = ow;
  *out_h = oh;
  return ow * oh;
}

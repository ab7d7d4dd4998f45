// =============================================================================
// pslam_oracle.hpp -- CPU ORACLE (TEST INFRASTRUCTURE, NOT THE PRODUCT)
//
// A plain C++17 restatement of the srrg2_proslam visual-odometry frontend hot
// path, used ONLY as the checker for the CUDA path (tests/, smoke(), and
// bench.py's cpu_baseline / --impl reference legs).  Nothing under
// srrg2_proslam_b200/ links, imports or calls this code.
//
// The reference cannot be compiled in this image (needs catkin, Eigen, OpenCV 3
// C++, srrg2_core, srrg2_solver, srrg2_slam_interfaces -- see DESIGN.md), so each
// function below restates the reference's algorithm and cites the file:line it
// follows (paths relative to the reference root, `srrg2_proslam/src/srrg2_proslam/`
// abbreviated as `.../`).  Third-party arithmetic that is not vendored in the
// reference is restated from its published algorithm:
//   * OpenCV 3.x (unpinned by the reference): cv::FastFeatureDetector (TYPE_9_16,
//     non-max suppression), cv::ORB::compute on provided keypoints (integer 7x7
//     Gaussian blur + bit_pattern_31_), cv::norm(NORM_HAMMING).
//   * libstdc++ std::sort / std::unordered_map: used DIRECTLY here (this file is
//     compiled with g++), so implementation-defined tie orders are the real ones.
//
// Parity pinning: stages 1-2 (detect, describe, match, adaptors) are pinned by the
// reference's own integer known answers (tests/test_oracle_known_answers.py, 26
// constants from the reference's gtest files).  Stages 3-4 (SE3 factors, H/b, GN)
// are "parity unpinned" at value level -- the reference tests only hold pose
// tolerances vs ground truth; see pslam_oracle_solver.hpp.
// =============================================================================
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <set>
#include <unordered_map>
#include <unordered_set>
#include <vector>

#include "../include/pslam_orb_pattern.h"

namespace pslam_oracle {

struct KeyPoint {  // the fields of cv::KeyPoint the path reads
  float x, y;
  float response;
};

struct Correspondence {  // srrg2_core::Correspondence{fixed_idx, moving_idx, response}
  int fixed_idx;
  int moving_idx;
  float response;
};
using CorrespondenceVector = std::vector<Correspondence>;

struct Descriptor {
  uint8_t b[32];
};

struct Feature2 {  // PointIntensityDescriptor_<D,float> restricted to what the path uses
  float x, y;      // coordinates()(0), coordinates()(1)
  float z = 0;     // coordinates()(2) where present (depth / right u)
  float w = 0;     // coordinates()(3) where present (right v)
  float intensity = 0;
  Descriptor desc;
};
using Cloud = std::vector<Feature2>;

// -----------------------------------------------------------------------------
// Hamming distance: srrg2_core descriptor field distance() == cv::norm(HAMMING)
// call sites: .../registration/correspondence_finders/
//   correspondence_finder_descriptor_based_bruteforce_impl.cpp:48-49,
//   correspondence_finder_descriptor_based_epipolar_impl.cpp:157,
//   correspondence_finder_projective_circle_impl.cpp:59-61
// -----------------------------------------------------------------------------
static inline int hamming256(const Descriptor& a, const Descriptor& b) {
  int d = 0;
  for (int i = 0; i < 32; i += 8) {
    uint64_t x, y;
    std::memcpy(&x, a.b + i, 8);
    std::memcpy(&y, b.b + i, 8);
    d += __builtin_popcountll(x ^ y);
  }
  return d;
}

// -----------------------------------------------------------------------------
// FAST-9/16 with optional 3x3 non-max suppression.
// Replaces cv::FastFeatureDetector::detect, called at
//   .../sensor_processing/feature_extractors/intensity_feature_extractor_binned.cpp:141-146
// created at .../intensity_feature_extractor_base.cpp:123-125 (threshold, nms; 9_16).
// Output is row-major (y then x), response = score (0 when nms is off).
// -----------------------------------------------------------------------------
static const int FAST_DX[16] = {0, 1, 2, 3, 3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1};
static const int FAST_DY[16] = {3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1, 0, 1, 2, 3};

// s = max over the 16 contiguous 9-arcs of min(d_k) and of min(-d_k); corner iff s > thr.
// min over 9 consecutive = min of three mins over 3 consecutive (k, k+3, k+6): 5 instead of 16 operations per arc.
// want_dark / want_bright: polarities worth evaluating (the other one contributes <= 0 when the compass pre-test
// excluded it: a 9-arc always contains two adjacent compass points); both = the full definition.
// Evaluated on the 16 ring pixels as one 16-byte vector (GCC vector extensions): saturating differences (a negative
// difference clips to 0, which cannot change a strength that exceeds a threshold >= 0 -- callers only compare the
// result with thr >= 0 or 0), min over 3 consecutive ring positions, then over three of those (9-arc), horizontal max.
typedef uint8_t vu8x16 __attribute__((vector_size(16)));
static inline vu8x16 vmin16(vu8x16 a, vu8x16 b) { return a < b ? a : b; }
static inline vu8x16 vmax16(vu8x16 a, vu8x16 b) { return a > b ? a : b; }
static inline int fast_arc_min9_max(vu8x16 d) {
  const vu8x16 r1 = __builtin_shufflevector(d, d, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 0);
  const vu8x16 r2 = __builtin_shufflevector(d, d, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 0, 1);
  const vu8x16 m3 = vmin16(d, vmin16(r1, r2));
  const vu8x16 s3 = __builtin_shufflevector(m3, m3, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 0, 1, 2);
  const vu8x16 s6 = __builtin_shufflevector(m3, m3, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 0, 1, 2, 3, 4, 5);
  vu8x16 m = vmin16(m3, vmin16(s3, s6));  // m[k] = min over ring positions k .. k + 8
  m = vmax16(m, __builtin_shufflevector(m, m, 8, 9, 10, 11, 12, 13, 14, 15, 0, 1, 2, 3, 4, 5, 6, 7));
  m = vmax16(m, __builtin_shufflevector(m, m, 4, 5, 6, 7, 0, 1, 2, 3, 4, 5, 6, 7, 0, 1, 2, 3));
  m = vmax16(m, __builtin_shufflevector(m, m, 2, 3, 0, 1, 2, 3, 0, 1, 2, 3, 0, 1, 2, 3, 0, 1));
  return std::max(m[0], m[1]);
}
static inline int fast_arc_strength(const uint8_t* p, int stride, bool want_dark = true, bool want_bright = true) {
  vu8x16 ring, c;
  for (int k = 0; k < 16; ++k) {
    ring[k] = p[FAST_DY[k] * stride + FAST_DX[k]];
    c[k] = p[0];
  }
  int s = 0;
  if (want_dark) s = std::max(s, fast_arc_min9_max(vmax16(c, ring) - ring));      // sat(c - ring)
  if (want_bright) s = std::max(s, fast_arc_min9_max(vmax16(ring, c) - c));       // sat(ring - c)
  return s;
}

// compass pre-test of 32 pixels at once (GCC vector extensions, bytes): ring > v + thr  <=>  sat(ring - v) > thr
typedef uint8_t vu8x32 __attribute__((vector_size(32)));
static inline vu8x32 vld32(const uint8_t* p) {
  vu8x32 v;
  std::memcpy(&v, p, 32);
  return v;
}
static inline vu8x32 vsatsub(vu8x32 a, vu8x32 b) {
  const vu8x32 m = a > b ? a : b;
  return m - b;
}

// score map: 0 for non-corners, s-1 (>= thr) for corners; interior pixels only.
static inline void fast_score_map(const uint8_t* img, int rows, int cols, int stride, int thr,
                                  std::vector<int>& score) {
  score.assign((size_t) rows * cols, 0);
  thr = std::min(std::max(thr, 0), 255);
  std::vector<uint8_t> flag((size_t) cols + 32, 0);
  vu8x32 T, one, two;
  for (int i = 0; i < 32; ++i) {
    T[i] = (uint8_t) thr;
    one[i] = 1;
    two[i] = 2;
  }
  for (int y = 3; y < rows - 3; ++y) {
    const uint8_t* row = img + (size_t) y * stride;
    const uint8_t *rn = row + 3 * stride, *rs = row - 3 * stride;
    // cheap reject: a 9-arc always contains two adjacent compass points -> (N | S) & (E | W) per polarity
    int x = 3;
    for (; x + 32 <= cols - 3; x += 32) {
      const vu8x32 v = vld32(row + x), n = vld32(rn + x), so = vld32(rs + x), e = vld32(row + x + 3), w = vld32(row + x - 3);
      const vu8x32 bright = (vu8x32) (((vsatsub(n, v) > T) | (vsatsub(so, v) > T)) & ((vsatsub(e, v) > T) | (vsatsub(w, v) > T)));
      const vu8x32 dark = (vu8x32) (((vsatsub(v, n) > T) | (vsatsub(v, so) > T)) & ((vsatsub(v, e) > T) | (vsatsub(v, w) > T)));
      const vu8x32 f = (bright & one) | (dark & two);
      std::memcpy(flag.data() + x, &f, 32);
    }
    for (; x < cols - 3; ++x) {
      const int v = row[x];
      const int c0 = rn[x], c4 = row[x + 3], c8 = rs[x], c12 = row[x - 3];
      const int hi = v + thr, lo = v - thr;
      const int bright = ((c0 > hi) | (c8 > hi)) & ((c4 > hi) | (c12 > hi));
      const int dark = ((c0 < lo) | (c8 < lo)) & ((c4 < lo) | (c12 < lo));
      flag[x] = (uint8_t) (bright | (dark << 1));
    }
    for (x = 3; x < cols - 3; ++x) {
      if (!flag[x]) continue;
      const int s = fast_arc_strength(row + x, stride, (flag[x] & 2) != 0, (flag[x] & 1) != 0);
      if (s > thr) score[(size_t) y * cols + x] = s - 1;
    }
  }
}

static inline void fast_detect(const uint8_t* img, int rows, int cols, int stride, int thr,
                               bool nms, const uint8_t* mask, int mask_stride,
                               std::vector<KeyPoint>& out) {
  out.clear();
  std::vector<int> score;
  fast_score_map(img, rows, cols, stride, thr, score);
  thr = std::min(std::max(thr, 0), 255);
  for (int y = 3; y < rows - 3; ++y) {
    for (int x = 3; x < cols - 3; ++x) {
      const int s = score[(size_t) y * cols + x];
      // a corner has s-1 >= thr; thr==0 corners can have score 0, so test the strength itself
      bool corner = s > 0;
      if (!corner && thr == 0) {
        corner = fast_arc_strength(img + (size_t) y * stride + x, stride) > 0;
      }
      if (!corner) continue;
      if (nms) {
        bool is_max = true;
        for (int dy = -1; dy <= 1 && is_max; ++dy)
          for (int dx = -1; dx <= 1; ++dx) {
            if (!dx && !dy) continue;
            if (!(s > score[(size_t)(y + dy) * cols + (x + dx)])) {
              is_max = false;
              break;
            }
          }
        if (!is_max) continue;
      }
      // masked detection == detect then drop (cv::KeyPointsFilter::runByPixelsMask)
      if (mask && mask[(size_t) y * mask_stride + x] == 0) continue;
      out.push_back(KeyPoint{(float) x, (float) y, nms ? (float) s : 0.0f});
    }
  }
}

// -----------------------------------------------------------------------------
// Binned extractor grid and selection.
// Follows IntensityFeatureExtractorBinned_::init
//   (.../feature_extractors/intensity_feature_extractor_binned.cpp:47-90) and
// ::computeKeypoints (:115-208).  Only the (r,c)->region LUT (:83-90) takes part in
// the selection; the cv::Rect regions (:47-71) are informational.
// -----------------------------------------------------------------------------
struct BinGrid {
  size_t rows = 0, cols = 0, nh = 0, nv = 0;
  size_t regions = 0;
  size_t quota = 0;  // _target_number_of_keypoints_per_detection_region
  float pixel_rows_per_detector = 0, pixel_cols_per_detector = 0;

  void init(size_t rows_, size_t cols_, size_t nh_, size_t nv_, int target) {
    rows = rows_;
    cols = cols_;
    nh = nh_;
    nv = nv_;
    regions = nv * nh;
    pixel_rows_per_detector = static_cast<float>(rows) / nv;  // :49-50
    pixel_cols_per_detector = static_cast<float>(cols) / nh;  // :51-52
    quota = static_cast<float>(target) / regions;             // :72-75 (float -> size_t)
  }
  // :83-90, evaluated with the same float operations and conversions
  size_t region(size_t r, size_t c) const {
    const size_t row_region = std::floor(r / pixel_rows_per_detector) * nh;
    const size_t idx = row_region + c / pixel_cols_per_detector;
    return idx;
  }
};

static inline void bin_select(const BinGrid& g, std::vector<KeyPoint>& keypoints) {
  // .../intensity_feature_extractor_binned.cpp:164-200
  std::vector<std::vector<KeyPoint>> per_region(g.regions);
  for (const KeyPoint& kp : keypoints) {
    const size_t r = kp.y;
    const size_t c = kp.x;
    per_region[g.region(r, c)].push_back(kp);
  }
  keypoints.clear();
  for (std::vector<KeyPoint>& avail : per_region) {
    if (avail.size() < g.quota) {
      keypoints.insert(keypoints.end(), avail.begin(), avail.end());
    } else {
      // unstable libstdc++ introsort decides which equal-response keypoints survive the cut
      std::sort(avail.begin(), avail.end(),
                [](const KeyPoint& a, const KeyPoint& b) { return a.response > b.response; });
      keypoints.insert(keypoints.end(), avail.begin(), avail.begin() + g.quota);
    }
  }
}

// -----------------------------------------------------------------------------
// ORB-256 on provided keypoints == cv::ORB::create()->compute(image, kps, desc)
// (call site .../intensity_feature_extractor_base.cpp:45-53; OpenCV 3.x arithmetic).
// -----------------------------------------------------------------------------
static const int ORB_EDGE = 31;
static const int BLUR_TAPS[7] = {18, 34, 49, 55, 49, 34, 18};  // round(256*gauss(sigma 2)), sum 257

static inline int reflect101(int i, int n) {
  if (n == 1) return 0;
  while (i < 0 || i >= n) {
    if (i < 0) i = -i;
    if (i >= n) i = 2 * (n - 1) - i;
  }
  return i;
}

// 7x7 sigma-2 Gaussian, BORDER_REFLECT_101, OpenCV-3 8-bit fixed point path:
// h = sum k*p (no rounding), v = sum k*h, out = sat_u8((v + 2^15) >> 16)
static inline void blur7(const uint8_t* img, int rows, int cols, int stride,
                         std::vector<uint8_t>& out) {
  out.assign((size_t) rows * cols, 0);
  // horizontal pass (h <= 257 * 255 = 65535 fits 16 bits); the interior runs without the border reflection
  std::vector<uint16_t> h((size_t) rows * cols);
  for (int y = 0; y < rows; ++y) {
    const uint8_t* r = img + (size_t) y * stride;
    uint16_t* hr = h.data() + (size_t) y * cols;
    for (int x = 0; x < cols; ++x) {
      if (x == 3 && cols > 6) {
        for (; x < cols - 3; ++x)
          hr[x] = (uint16_t) (18 * (r[x - 3] + r[x + 3]) + 34 * (r[x - 2] + r[x + 2]) + 49 * (r[x - 1] + r[x + 1]) + 55 * r[x]);
        if (x >= cols) break;
      }
      int acc = 0;
      for (int j = -3; j <= 3; ++j) acc += BLUR_TAPS[j + 3] * r[reflect101(x + j, cols)];
      hr[x] = (uint16_t) acc;
    }
  }
  // vertical pass, x innermost (contiguous)
  for (int y = 0; y < rows; ++y) {
    const uint16_t* hp[7];
    for (int j = -3; j <= 3; ++j) hp[j + 3] = h.data() + (size_t) reflect101(y + j, rows) * cols;
    uint8_t* o = out.data() + (size_t) y * cols;
    for (int x = 0; x < cols; ++x) {
      const int acc = 18 * (hp[0][x] + hp[6][x]) + 34 * (hp[1][x] + hp[5][x]) + 49 * (hp[2][x] + hp[4][x]) + 55 * hp[3][x];
      o[x] = (uint8_t) std::min(255, (acc + 32768) >> 16);
    }
  }
}

static inline void orb_border_filter(int rows, int cols, std::vector<KeyPoint>& kps) {
  std::vector<KeyPoint> kept;
  kept.reserve(kps.size());
  for (const KeyPoint& k : kps) {
    if (k.x >= ORB_EDGE && k.x < cols - ORB_EDGE && k.y >= ORB_EDGE && k.y < rows - ORB_EDGE) {
      kept.push_back(k);
    }
  }
  kps.swap(kept);
}

static inline Descriptor orb_describe(const uint8_t* blurred, int cols, int x, int y) {
  Descriptor d;
  for (int i = 0; i < 32; ++i) {
    int byte = 0;
    for (int k = 0; k < 8; ++k) {
      const signed char* p = PSLAM_ORB_PATTERN + 4 * (8 * i + k);
      const int a = blurred[(size_t)(y + p[1]) * cols + (x + p[0])];
      const int b = blurred[(size_t)(y + p[3]) * cols + (x + p[2])];
      byte |= (a < b) << k;
    }
    d.b[i] = (uint8_t) byte;
  }
  return d;
}

// -----------------------------------------------------------------------------
// IntensityFeatureExtractorBinned_::compute(cv::Mat)
//   == IntensityFeatureExtractor_::compute (.../intensity_feature_extractor_base.cpp:55-85)
// -----------------------------------------------------------------------------
struct ExtractConfig {
  float detector_threshold = 10;   // PARAM detector_threshold  (base.h:36-40)
  int enable_nms = 1;              // PARAM enable_non_maximum_suppression (:48-52)
  int target_number_of_keypoints = 500;  // (:54-58)
  int detectors_horizontal = 3;    // binned.h:17-21
  int detectors_vertical = 3;      // binned.h:23-27
};

static inline void extract_binned(const uint8_t* img, int rows, int cols, int stride,
                                  const ExtractConfig& cfg, const uint8_t* mask, int mask_stride,
                                  Cloud& features, std::vector<float>* responses = nullptr) {
  features.clear();
  BinGrid grid;
  grid.init(rows, cols, cfg.detectors_horizontal, cfg.detectors_vertical,
            cfg.target_number_of_keypoints);
  std::vector<KeyPoint> kps;
  fast_detect(img, rows, cols, stride, (int) cfg.detector_threshold, cfg.enable_nms != 0, mask,
              mask_stride, kps);
  if (!mask) bin_select(grid, kps);  // binned.cpp:167 "only perform binning if mask is not set"
  orb_border_filter(rows, cols, kps);
  std::vector<uint8_t> blurred;
  blur7(img, rows, cols, stride, blurred);
  features.reserve(kps.size());
  if (responses) responses->clear();
  for (const KeyPoint& k : kps) {
    Feature2 f;
    f.x = k.x;
    f.y = k.y;
    f.intensity = img[(size_t) k.y * stride + (size_t) k.x];
    f.desc = orb_describe(blurred.data(), cols, (int) k.x, (int) k.y);
    features.push_back(f);
    if (responses) responses->push_back(k.response);
  }
}

// -----------------------------------------------------------------------------
// CorrespondenceFinderDescriptorBasedEpipolar::compute
//   .../correspondence_finders/correspondence_finder_descriptor_based_epipolar_impl.cpp:44-219
// -----------------------------------------------------------------------------
struct EpipolarConfig {
  float maximum_descriptor_distance = 50.0f;            // bruteforce.h:23-27
  float maximum_distance_ratio_to_second_best = 0.9f;   // :28-32
  unsigned maximum_disparity_pixels = 100;              // epipolar.h:24-28
  unsigned epipolar_line_thickness_pixels = 0;          // :30-34
};

struct SortedFeature {  // epipolar_impl.cpp:8-23
  int32_t row, col, unsorted_index;
};

static inline void sort_feature_vector(const Cloud& cloud, std::vector<SortedFeature>& v) {
  v.clear();
  v.reserve(cloud.size());
  for (size_t i = 0; i < cloud.size(); ++i) {
    v.push_back(SortedFeature{(int32_t) cloud[i].y, (int32_t) cloud[i].x, (int32_t) i});
  }
  std::sort(v.begin(), v.end(), [](const SortedFeature& a, const SortedFeature& b) {
    return (a.row < b.row) || (a.row == b.row && a.col < b.col);
  });
}

static inline void match_epipolar(const Cloud& fixed, const Cloud& moving,
                                  const EpipolarConfig& cfg, CorrespondenceVector& out) {
  out.clear();
  const float max_dist = cfg.maximum_descriptor_distance;
  const float max_ratio = cfg.maximum_distance_ratio_to_second_best;
  const int32_t max_disp = cfg.maximum_disparity_pixels;
  std::vector<SortedFeature> L, R;
  sort_feature_vector(fixed, L);
  sort_feature_vector(moving, R);
  std::vector<int32_t> row_offsets{0};
  for (int32_t o = 1; o < 1 + (int32_t) cfg.epipolar_line_thickness_pixels; ++o) {
    row_offsets.push_back(o);
    row_offsets.push_back(-o);
  }
  for (const int32_t off : row_offsets) {
    uint32_t index_right = 0;
    std::set<size_t> matched_left, matched_right;
    if (R.empty()) break;  // reference would read R[0] of an empty vector; nothing can match
    for (size_t index_left = 0; index_left < L.size(); ++index_left) {
      if (index_right == R.size()) break;
      while (L[index_left].row + off < R[index_right].row) {
        ++index_left;
        if (index_left == L.size()) break;
      }
      if (index_left == L.size()) break;
      const int row_left = L[index_left].row + off;
      const int col_left = L[index_left].col;
      const size_t unsorted_left = L[index_left].unsorted_index;
      const Descriptor& dl = fixed[unsorted_left].desc;
      while (row_left > R[index_right].row) {
        ++index_right;
        if (index_right == R.size()) break;
      }
      if (index_right == R.size()) break;
      size_t s = index_right;
      float best = std::numeric_limits<float>::max();
      float second = std::numeric_limits<float>::max();
      size_t best_right = 0;
      // the reference reads R[s] before the bounds test (:136-137); the observable
      // behaviour is "stop at the end", restated with the test first.
      while (s < R.size() && row_left == R[s].row) {
        const int32_t disparity = col_left - R[s].col;
        if (disparity < 0) break;
        if (disparity > max_disp) {
          ++s;
          continue;
        }
        const int d = hamming256(dl, moving[R[s].unsorted_index].desc);
        if (d < best) {
          second = best;
          best = d;
          best_right = s;
        } else if (d < second) {
          second = d;
        }
        ++s;
      }
      if (best < max_dist && best / second < max_ratio) {
        out.push_back(Correspondence{(int) unsorted_left, R[best_right].unsorted_index, best});
        index_right = best_right + 1;
        matched_left.insert(index_left);
        matched_right.insert(best_right);
      }
    }
    size_t keep = 0;
    for (size_t i = 0; i < L.size(); ++i)
      if (!matched_left.count(i)) L[keep++] = L[i];
    L.resize(keep);
    keep = 0;
    for (size_t i = 0; i < R.size(); ++i)
      if (!matched_right.count(i)) R[keep++] = R[i];
    R.resize(keep);
  }
}

// -----------------------------------------------------------------------------
// CorrespondenceFinderDescriptorBasedBruteforce::compute (+ checkLowesRatio,
// _processCorrespondencePool)
//   .../correspondence_finders/correspondence_finder_descriptor_based_bruteforce_impl.cpp:6-294
// -----------------------------------------------------------------------------
struct BruteforceConfig {
  float maximum_descriptor_distance = 50.0f;
  float maximum_distance_ratio_to_second_best = 0.9f;
};

static inline bool lowes_pair(float best, float other, float max_ratio) {  // :157-176
  if (best == other) return false;
  return best / other < max_ratio;
}
static inline bool lowes_list(float best, const std::vector<float>& sorted, float max_ratio) {
  if (sorted.size() == 1) return true;  // :185-188
  float second = best;
  for (float d : sorted)
    if (d > best) {
      second = d;
      break;
    }
  return lowes_pair(best, second, max_ratio);
}

static inline void bf_process_pool(const CorrespondenceVector& pool,
                                   std::unordered_map<int, std::vector<float>>& dist_fixed,
                                   std::unordered_map<int, std::vector<float>>& dist_moving,
                                   std::unordered_set<int>& reg_fixed,
                                   std::unordered_set<int>& reg_moving, float max_ratio,
                                   CorrespondenceVector& out) {  // :245-294
  for (size_t j = 0; j < pool.size(); ++j) {
    bool unique = true;
    for (size_t k = 0; k < pool.size(); ++k)
      if (j != k && (pool[j].fixed_idx == pool[k].fixed_idx ||
                     pool[j].moving_idx == pool[k].moving_idx))
        unique = false;
    if (!unique) continue;
    const Correspondence& c = pool[j];
    if (lowes_list(c.response, dist_fixed[c.fixed_idx], max_ratio) &&
        lowes_list(c.response, dist_moving[c.moving_idx], max_ratio)) {
      out.push_back(c);
      reg_fixed.insert(c.fixed_idx);
      reg_moving.insert(c.moving_idx);
    }
  }
}

static inline void match_bruteforce(const Cloud& fixed, const Cloud& moving,
                                    const BruteforceConfig& cfg, CorrespondenceVector& out) {
  out.clear();
  const size_t nf = fixed.size(), nm = moving.size();
  const float max_dist = cfg.maximum_descriptor_distance;
  CorrespondenceVector cand;
  std::unordered_map<int, std::vector<float>> dist_fixed, dist_moving;
  dist_fixed.reserve(nf);
  dist_moving.reserve(nm);
  for (size_t f = 0; f < nf; ++f) {
    auto itf = dist_fixed.insert(std::make_pair((int) f, std::vector<float>())).first;
    for (size_t m = 0; m < nm; ++m) {
      const float d = hamming256(fixed[f].desc, moving[m].desc);
      if (d < max_dist) {
        cand.push_back(Correspondence{(int) f, (int) m, d});
        itf->second.push_back(d);
        dist_moving[(int) m].push_back(d);
      }
    }
    std::sort(itf->second.begin(), itf->second.end());
  }
  if (cand.empty()) return;
  if (cand.size() == 1) {
    out.push_back(cand.back());
    return;
  }
  std::sort(cand.begin(), cand.end(),
            [](const Correspondence& a, const Correspondence& b) { return a.response < b.response; });
  for (auto& kv : dist_moving) std::sort(kv.second.begin(), kv.second.end());
  std::unordered_set<int> reg_fixed, reg_moving;
  CorrespondenceVector pool(1, cand.front());
  for (size_t i = 1; i < cand.size(); ++i) {
    const Correspondence& c = cand[i];
    if (!reg_fixed.count(c.fixed_idx) && !reg_moving.count(c.moving_idx)) {
      // reference compares against pool.back() even when the pool is empty (:115); either
      // branch then ends with the candidate alone in the pool -- restated without the UB.
      if (!pool.empty() && c.response == pool.back().response) {
        pool.push_back(c);
      } else {
        bf_process_pool(pool, dist_fixed, dist_moving, reg_fixed, reg_moving,
                        cfg.maximum_distance_ratio_to_second_best, out);
        pool.clear();
        if (!reg_fixed.count(c.fixed_idx) && !reg_moving.count(c.moving_idx)) pool.push_back(c);
      }
    }
    if (reg_fixed.size() == nf || reg_moving.size() == nm) break;
  }
  if (!pool.empty())
    bf_process_pool(pool, dist_fixed, dist_moving, reg_fixed, reg_moving,
                    cfg.maximum_distance_ratio_to_second_best, out);
}

// -----------------------------------------------------------------------------
// RawDataPreprocessorStereoProjective::compute
//   .../sensor_processing/raw_data_preprocessor_stereo_projective.cpp:46-134
// stereo_matches -> (uL, vL, uR, vR) + left descriptor/intensity; negative disparities dropped
// -----------------------------------------------------------------------------
static inline void assemble_stereo_points(const Cloud& left, const Cloud& right,
                                          const CorrespondenceVector& matches, Cloud& meas) {
  meas.clear();
  meas.reserve(matches.size());
  for (const Correspondence& m : matches) {
    Feature2 p;
    p.x = left[m.fixed_idx].x;
    p.y = left[m.fixed_idx].y;
    p.z = right[m.moving_idx].x;
    p.w = right[m.moving_idx].y;
    p.desc = left[m.fixed_idx].desc;
    p.intensity = left[m.fixed_idx].intensity;
    const float hd = p.x - p.z, vd = p.y - p.w;
    if (hd < 0 || vd < 0) continue;  // :120-128
    meas.push_back(p);
  }
}

// -----------------------------------------------------------------------------
// RawDataPreprocessorMonocularDepth::_readDepth
//   .../sensor_processing/raw_data_preprocessor_monocular_depth.cpp:156-180
// -----------------------------------------------------------------------------
template <typename DepthT>
static inline void read_depth(Cloud& meas, const DepthT* depth, int depth_stride, float scale) {
  size_t keep = 0;
  for (size_t i = 0; i < meas.size(); ++i) {
    Feature2 f = meas[i];
    const float d = depth[(size_t) std::rint(f.y) * depth_stride + (size_t) std::rint(f.x)];
    if (d > 0) {
      f.z = scale * d;
      meas[keep++] = f;
    }
  }
  meas.resize(keep);
}

}  // namespace pslam_oracle

// k_fast.cu -- stage 1 front half (sm_100a): FAST-9/16 + 3x3 NMS fused with the ORB 7x7 integer blur in
// ONE pass over the image, and the per-region ordered gather + std::sort-exact selection.
//
// Reference path replaced:
//   cv::FastFeatureDetector::detect (TYPE_9_16) called at
//     .../sensor_processing/feature_extractors/intensity_feature_extractor_binned.cpp:139-147
//   the 7x7 sigma-2 Gaussian inside cv::ORB::compute (OpenCV 3 fixed point), called at
//     .../feature_extractors/intensity_feature_extractor_base.cpp:52
//   bucket + std::sort + quota of IntensityFeatureExtractorBinned_::computeKeypoints (binned.cpp:164-200)
//
// K1 `fast_blur_rows_kernel` -- "marching" design.  A CTA owns a band of bh image rows over the whole image
// width; a warp owns a 256-pixel strip, a lane 8 adjacent pixels (two 32-bit words).  The warp walks down
// the band one row per step and keeps a 7-row window of (a) the raw pixels and (b) the horizontally
// filtered rows in REGISTERS, so every pixel is loaded from HBM/L2 once and the inner loops are word-wide:
//   * blur: horizontal 7-tap as 2 x dp4a on byte quads, vertical 7-tap as 7 x dp2a on packed u16 pairs
//     (exact: sum k_i k_j p <= 257*257*255 < 2^25), (v + 2^15) >> 16, min 255;
//   * FAST compass pre-test on 16-bit SWAR pixel pairs: ring > c + t  <=>  bit 9 of (r + 512 - c - t - 1);
//     a 9-arc always contains two adjacent compass points, so (N|S) & (E|W) per polarity is necessary;
//   * candidates go to a per-warp shared-memory queue and are scored 32 at a time, one per lane, with only
//     the polarity the pre-test saw (a pixel cannot be a bright AND a dark 9-arc corner: 9 + 9 > 16);
//     score = max over the 16 arcs of min over the arc, written to the warp's own 16-row score RING;
//   * ROLLING 3x3 NMS inside the march: a fixed number of rows behind the scoring front the warp takes the three
//     ring rows it needs, computes the 8-neighbour maximum of its 8 pixels per lane with byte-SIMD max / compare,
//     and appends the survivors (col, response + 1) to the row's keypoint list in column order.
// Strips start every 252 pixels and cover 256: the 2 pixels of overlap on either side make a strip's NMS
// self-contained, so the warps of a CTA never synchronise, no score tile of the whole band exists (it capped the
// band height at ~29 rows = 28 % of re-marched halo rows and the CTAs at four per SM) and the band height is free.
// Every (row, strip) has its own 256-entry keypoint list; row-major order = rows, then strips, then entries = the
// order cv::FAST emits.  No dense NMS map ever goes to HBM.
// Blur values closer than 3 px to the image border are NOT the reflect-101 values (nothing on the path reads
// them: ORB only describes keypoints >= 31 px from the border and samples <= 13 px + the 3 px filter reach);
// `blur_border_kernel` patches them for the stage-level entry point pslam_blur7.
//
// K2 `bin_select_kernel`: per (region, image) gather of the region's keypoints from the row lists in
// row-major order, then the reference's selection (keep all if fewer than quota, else unstable std::sort by
// response and keep the first quota) replayed move-for-move (libstdcxx_sort.h).
#include <mutex>
#include <vector>

#include "libstdcxx_sort.h"
#include "pslam_internal.cuh"
#include "pslam_kernels.cuh"

namespace {

constexpr int QCAP = 1024;        // per-warp candidate ring (16-bit entries = 2 KB); one push phase adds <= 512
constexpr int STRIP_STRIDE = 252; // strips start every 252 px and cover 256 (emit: strip columns 2 .. 253; strip 0 from column 0)
constexpr int STRIP_LIST = 256;   // keypoint list entries per (row, strip)
constexpr int RING = 16;          // rows of the per-warp score ring (power of two)
constexpr int WARP_SMEM = RING * 256 + QCAP * 2 + RING * 4;  // score ring | candidate queue | queue tail after each row
constexpr unsigned FULL = 0xffffffffu;
constexpr unsigned M16 = 0x00ff00ffu;
constexpr unsigned K9 = 0x02000200u;  // bit 9 of both 16-bit lanes

__host__ __device__ inline int k1_strips(int cols) { return cols <= 256 ? 1 : (cols - 2 + STRIP_STRIDE - 1) / STRIP_STRIDE; }

// ---- row load: every lane fetches ONE naturally aligned 8-byte chunk ------------------------------------------
// Image rows start at any byte (KITTI: 1241-byte pitch).  With a = row + warp * 256 + lane * 8 and d = a & 7 (the same
// for every lane and warp of a row) the lane loads the chunk at a - d; its 8 pixels are bytes d .. d + 7 of (own chunk,
// next lane's chunk), the next chunk comes by shuffle.  Only the strip-edge lanes load more: lane 0 the chunk before
// (left halo), lane 31 the two chunks after.  Aligned chunks never straddle an allocation granule, so a chunk that
// starts inside the image is always readable; chunks that would start behind the row's last pixel are clamped to the
// chunk holding that pixel (their bytes belong to columns >= cols, which nothing uses).  Split in two so that the loads
// of row r + 1 are in flight while row r is processed.
struct RawQ {
  uint2 q, e0, e1;
};
// per-lane byte offsets of the chunks relative to (the strip's first aligned chunk - 8): constant over the march
struct LaneOffs {
  unsigned q, e0, e1;
  bool p_e0, p_e1;
};
__device__ __forceinline__ LaneOffs lane_offs(int lane, bool has_left) {
  LaneOffs o;
  o.q = 8u + 8u * (unsigned) lane;
  o.e0 = lane == 31 ? 8u + 256u : 0u;  // lane 0: the chunk before its own; lane 31: the chunk after
  o.e1 = 8u + 264u;
  o.p_e0 = (lane == 0 && has_left) || lane == 31;
  o.p_e1 = lane == 31;
  return o;
}
__device__ __forceinline__ RawQ loadq_issue(const uint8_t* __restrict__ row, int xs, const LaneOffs& lo, int cols) {
  // uniform over the warp: the aligned chunk that holds the strip's first pixel (minus one chunk) and the offset of the
  // chunk that holds the row's last pixel
  const uint8_t* a = row + xs;
  const uint8_t* base = a - ((uintptr_t) a & 7u) - 8;
  const uint8_t* last = row + (cols - 1);
  last -= (uintptr_t) last & 7u;
  const unsigned last_off = (unsigned) (last - base);
  RawQ r;
  r.e0 = r.e1 = make_uint2(0u, 0u);
  r.q = __ldg(reinterpret_cast<const uint2*>(base + min(lo.q, last_off)));
  if (lo.p_e0) r.e0 = __ldg(reinterpret_cast<const uint2*>(base + min(lo.e0, last_off)));
  if (lo.p_e1) r.e1 = __ldg(reinterpret_cast<const uint2*>(base + min(lo.e1, last_off)));
  return r;
}
// the 16-pixel window of a lane (4 px left halo | 8 own | 4 px right halo) from the chunks; d = address of the strip's first
// pixel & 7 (uniform)
struct RowRegs {
  unsigned hl, a0, a1, hr;  // pixels x0-4..x0-1 | x0..x0+3 | x0+4..x0+7 | x0+8..x0+11
};
__device__ __forceinline__ RowRegs assemble_row(const RawQ& nx, unsigned d, int lane) {
  const bool hi = d >= 4u;
  const unsigned sh = (d & 3u) * 8u;
  uint2 qn;
  qn.x = __shfl_down_sync(FULL, nx.q.x, 1);
  qn.y = __shfl_down_sync(FULL, nx.q.y, 1);
  if (lane == 31) qn = nx.e0;
  const unsigned w0 = hi ? nx.q.y : nx.q.x, w1 = hi ? qn.x : nx.q.y, w2 = hi ? qn.y : qn.x;
  RowRegs cur;
  cur.a0 = __funnelshift_r(w0, w1, sh);
  cur.a1 = __funnelshift_r(w1, w2, sh);
  cur.hl = __shfl_up_sync(FULL, cur.a1, 1);
  cur.hr = __shfl_down_sync(FULL, cur.a0, 1);
  if (lane == 0) cur.hl = hi ? __funnelshift_r(nx.q.x, nx.q.y, sh) : __funnelshift_r(nx.e0.y, nx.q.x, sh);
  if (lane == 31) cur.hr = hi ? __funnelshift_r(nx.e0.y, nx.e1.x, sh) : __funnelshift_r(nx.e0.x, nx.e0.y, sh);
  return cur;
}

// ---- FAST score of one candidate, one polarity (bright: ring brighter than the centre) ---------------------
// 16-bit SIMD: ring differences d_k = +-(c - r_k) biased by 256 (1 .. 511), d_k in the low and d_(k+8) in the high
// half of one register; min over 3 consecutive, then over 3 of those (9-arc), max over the 16 arcs: 8 + 8 three-input
// min and 4 max instructions instead of 80 scalar ones.
__device__ __forceinline__ unsigned swap16(unsigned x) { return __byte_perm(x, 0u, 0x1032); }
__device__ __forceinline__ int fast_score_polar(const uint8_t* __restrict__ p, int stride, bool bright) {
  const unsigned c = __ldg(p);
  const uint8_t *pm3 = p - 3 * stride, *pm2 = p - 2 * stride, *pm1 = p - stride, *pp1 = p + stride, *pp2 = p + 2 * stride,
                *pp3 = p + 3 * stride;
  unsigned r[16];
  r[0] = __ldg(pp3);
  r[1] = __ldg(pp3 + 1);
  r[2] = __ldg(pp2 + 2);
  r[3] = __ldg(pp1 + 3);
  r[4] = __ldg(p + 3);
  r[5] = __ldg(pm1 + 3);
  r[6] = __ldg(pm2 + 2);
  r[7] = __ldg(pm3 + 1);
  r[8] = __ldg(pm3);
  r[9] = __ldg(pm3 - 1);
  r[10] = __ldg(pm2 - 2);
  r[11] = __ldg(pm1 - 3);
  r[12] = __ldg(p - 3);
  r[13] = __ldg(pp1 - 3);
  r[14] = __ldg(pp2 - 2);
  r[15] = __ldg(pp3 - 1);
  // bright: d = r - c, dark: d = c - r = (255 - r) - (255 - c); + 256 per half (every half stays in 1 .. 511, no borrow
  // between the halves).  The polarity is one XOR on the packed ring pairs instead of a multiply / negate per pair.
  const unsigned px = bright ? 0u : 0x00ff00ffu;
  const unsigned kb = (256u - (c ^ (px & 0xffu))) * 0x00010001u;
  unsigned d[10];
#pragma unroll
  for (int k = 0; k < 8; ++k) d[k] = kb + (__byte_perm(r[k], r[k + 8], 0x5410) ^ px);
  d[8] = swap16(d[0]);
  d[9] = swap16(d[1]);
  unsigned lo3[14];
#pragma unroll
  for (int k = 0; k < 8; ++k) lo3[k] = __vimin3_u16x2(d[k], d[k + 1], d[k + 2]);
#pragma unroll
  for (int k = 0; k < 6; ++k) lo3[8 + k] = swap16(lo3[k]);
  unsigned a[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) a[k] = __vimin3_u16x2(lo3[k], lo3[k + 3], lo3[k + 6]);
  unsigned m = __vimax3_u16x2(a[0], a[1], a[2]);
  m = __vimax3_u16x2(m, a[3], a[4]);
  m = __vimax3_u16x2(m, a[5], a[6]);
  m = __vmaxu2(m, a[7]);
  return (int) max(m & 0xffffu, m >> 16) - 256;
}

// horizontal 7-tap of the 8 own pixels -> 8 sums (<= 257 * 255 = 65535 each)
__device__ __forceinline__ void hblur8(const RowRegs& r, unsigned h[8]) {
  const unsigned KA = 18u | (34u << 8) | (49u << 16) | (55u << 24);  // taps -3..0
  const unsigned KB = 49u | (34u << 8) | (18u << 16);                // taps +1..+3
  // byte quads of the 16-byte window at offsets 1..12 (offset o = pixel x0 - 4 + o)
  const unsigned q1 = __byte_perm(r.hl, r.a0, 0x4321), q2 = __byte_perm(r.hl, r.a0, 0x5432),
                 q3 = __byte_perm(r.hl, r.a0, 0x6543), q4 = r.a0;
  const unsigned q5 = __byte_perm(r.a0, r.a1, 0x4321), q6 = __byte_perm(r.a0, r.a1, 0x5432),
                 q7 = __byte_perm(r.a0, r.a1, 0x6543), q8 = r.a1;
  const unsigned q9 = __byte_perm(r.a1, r.hr, 0x4321), q10 = __byte_perm(r.a1, r.hr, 0x5432),
                 q11 = __byte_perm(r.a1, r.hr, 0x6543), q12 = r.hr;
  h[0] = __dp4a(q1, KA, __dp4a(q5, KB, 0u));
  h[1] = __dp4a(q2, KA, __dp4a(q6, KB, 0u));
  h[2] = __dp4a(q3, KA, __dp4a(q7, KB, 0u));
  h[3] = __dp4a(q4, KA, __dp4a(q8, KB, 0u));
  h[4] = __dp4a(q5, KA, __dp4a(q9, KB, 0u));
  h[5] = __dp4a(q6, KA, __dp4a(q10, KB, 0u));
  h[6] = __dp4a(q7, KA, __dp4a(q11, KB, 0u));
  h[7] = __dp4a(q8, KA, __dp4a(q12, KB, 0u));
}

// vertical 7-tap on ROW-PAIR packed words: p_k = h(row 2k, x) | h(row 2k + 1, x) << 16 for the four row pairs of the 8-row
// window (rows w .. w + 7).  One dp2a = two taps: 4 instead of 7 multiply-accumulates per output pixel.
//   first  output row (centre w + 3, rows w .. w + 6):     taps (18,34) (49,55) (49,34) (18, 0)  = the low bytes of T0..T3
//   second output row (centre w + 4, rows w + 1 .. w + 7): taps ( 0,18) (34,49) (55,49) (34,18)  = the high bytes
// exact: the sum is <= 257 * 65535 + 2^15 < 2^25; (v + 2^15) >> 16, saturated to 255 (byte 2 after min with 0xffffff)
__device__ __forceinline__ void vblur_pairs(unsigned p0, unsigned p1, unsigned p2, unsigned p3, unsigned& first, unsigned& second) {
  const unsigned T0 = 18u | (34u << 8) | (0u << 16) | (18u << 24), T1 = 49u | (55u << 8) | (34u << 16) | (49u << 24);
  const unsigned T2 = 49u | (34u << 8) | (55u << 16) | (49u << 24), T3 = 18u | (0u << 8) | (34u << 16) | (18u << 24);
  unsigned f = __dp2a_lo(p0, T0, 32768u), g = __dp2a_hi(p0, T0, 32768u);
  f = __dp2a_lo(p1, T1, f); g = __dp2a_hi(p1, T1, g);
  f = __dp2a_lo(p2, T2, f); g = __dp2a_hi(p2, T2, g);
  f = __dp2a_lo(p3, T3, f); g = __dp2a_hi(p3, T3, g);
  first = min(f, 0x00ffffffu);
  second = min(g, 0x00ffffffu);
}
// bytes 2 of four sums -> one word of four output pixels
__device__ __forceinline__ unsigned pack_b2(unsigned v0, unsigned v1, unsigned v2, unsigned v3) {
  return __byte_perm(__byte_perm(v0, v1, 0x0062), __byte_perm(v2, v3, 0x0062), 0x5410);
}

// compass pre-test for the 8 own pixels of the centre row; returns 8-bit masks (bit j = pixel x0 + j)
struct RowOwn {
  unsigned a0, a1;  // pixels x0..x0+3 | x0+4..x0+7
};
__device__ __forceinline__ void pretest8(const RowOwn& n, const RowRegs& c, const RowOwn& s, unsigned t1,
                                         unsigned& m_bright, unsigned& m_dark) {
  // centre pairs: E0 = (p0,p2) O0 = (p1,p3) E1 = (p4,p6) O1 = (p5,p7)
  const unsigned cE0 = c.a0 & M16, cO0 = __byte_perm(c.a0, 0u, 0x4341), cE1 = c.a1 & M16, cO1 = __byte_perm(c.a1, 0u, 0x4341);
  const unsigned nE0 = n.a0 & M16, nO0 = __byte_perm(n.a0, 0u, 0x4341), nE1 = n.a1 & M16, nO1 = __byte_perm(n.a1, 0u, 0x4341);
  const unsigned sE0 = s.a0 & M16, sO0 = __byte_perm(s.a0, 0u, 0x4341), sE1 = s.a1 & M16, sO1 = __byte_perm(s.a1, 0u, 0x4341);
  // east (+3) / west (-3) ring pairs on the centre row
  const unsigned f36 = __funnelshift_r(c.a0, c.a1, 24);   // p3..p6
  const unsigned f710 = __funnelshift_r(c.a1, c.hr, 24);  // p7..p10
  const unsigned gm30 = __funnelshift_r(c.hl, c.a0, 8);   // p-3..p0
  const unsigned f25 = __funnelshift_r(c.a0, c.a1, 16);   // p2..p5
  const unsigned eE0 = f36 & M16 /*(3,5)*/, eO0 = cE1 /*(4,6)*/, eE1 = f710 & M16 /*(7,9)*/, eO1 = c.hr & M16 /*(8,10)*/;
  const unsigned wE0 = gm30 & M16 /*(-3,-1)*/, wO0 = __byte_perm(gm30, 0u, 0x4341) /*(-2,0)*/, wE1 = cO0 /*(1,3)*/,
                 wO1 = f25 & M16 /*(2,4)*/;
#define PSLAM_PRETEST(C, N, S, E, W, OUTB, OUTD)                                            \
  {                                                                                         \
    const unsigned kb = K9 - (C) - t1; /* + ring: bit 9 <=> ring > c + t */                 \
    const unsigned kd = K9 + (C) - t1; /* - ring: bit 9 <=> ring < c - t */                 \
    OUTB = (((N) + kb) | ((S) + kb)) & (((E) + kb) | ((W) + kb)) & K9;                      \
    OUTD = ((kd - (N)) | (kd - (S))) & ((kd - (E)) | (kd - (W))) & K9;                      \
  }
  unsigned bE0, bO0, bE1, bO1, dE0, dO0, dE1, dO1;
  PSLAM_PRETEST(cE0, nE0, sE0, eE0, wE0, bE0, dE0)
  PSLAM_PRETEST(cO0, nO0, sO0, eO0, wO0, bO0, dO0)
  PSLAM_PRETEST(cE1, nE1, sE1, eE1, wE1, bE1, dE1)
  PSLAM_PRETEST(cO1, nO1, sO1, eO1, wO1, bO1, dO1)
#undef PSLAM_PRETEST
  // bit 9 -> pixel (0|1|4|5), bit 25 -> pixel (2|3|6|7)
  const unsigned ub = (bE0 >> 9) | (bO0 >> 8) | (bE1 >> 5) | (bO1 >> 4);
  const unsigned ud = (dE0 >> 9) | (dO0 >> 8) | (dE1 >> 5) | (dO1 >> 4);
  m_bright = (ub | (ub >> 14)) & 0xffu;
  m_dark = (ud | (ud >> 14)) & 0xffu;
}

struct K1Args {
  const uint8_t* images;
  long long image_pitch;
  int rows, cols, stride, thr, nms;
  int bh;              // rows per band, chosen by the launcher
  uint8_t* blur;
  int map_pitch;
  long long map_slot;
  int* row_count;      // [images][max_rows][strips_cap]
  uint32_t* row_kp;    // [images][max_rows][strips_cap][STRIP_LIST]  (col << 8) | (response + 1)
  int strips_cap, max_rows, n_strips;
};

// Candidate queue: push this lane's candidates (mask bits 0-7 bright, 8-15 dark pixels x0..x0+7 of tile row
// trow) in lane order, then score full batches of 32, one candidate per lane.  Deliberately NOT inlined: it is
// called from the marching loop and inlining it makes the kernel overflow the instruction cache
// (ncu: stall_no_instruction was the top stall reason).  Returns the new (head, tail) packed in 64 bits.
// Entries are 16 bits, see score_entry.
struct ScoreCtx {
  const uint8_t* img;   // first pixel of the strip (image + strip * STRIP_STRIDE)
  uint8_t* s_score;     // the warp's score ring: [RING][256], strip column k at byte k
  int stride, thr, row_lo;
};
__device__ __forceinline__ void score_entry(const ScoreCtx& c, unsigned entry) {
  // entry = [ring row of the pair's first row : 4][lane : 5][bit index i : 5], i = [second row : 1][dark : 1][pixel : 3] -- the
  // push loop (one or two active lanes per iteration) only adds i, the decoding happens here with all 32 lanes busy.
  // Pending entries belong to rows row_lo .. row_lo + RING - 1: the ring row identifies the image row.
  const unsigned i = entry & 31u;
  const int col = (int) ((entry >> 2) & 0xf8u) + (int) (i & 7u);
  const int rr = (int) ((entry >> 10) + (i >> 4)) & (RING - 1);
  const bool bright = (i & 8u) == 0u;
  const int y = c.row_lo + ((rr - c.row_lo) & (RING - 1));
  const int s = fast_score_polar(c.img + (size_t) y * c.stride + col, c.stride, bright);
  if (s > c.thr) c.s_score[rr * 256 + col] = (uint8_t) s;  // 1..255; response = s - 1
}
// One call per PAIR of rows (A, A + 1), three jobs (deliberately NOT inlined: inlining it into the marching loop makes
// the kernel overflow the instruction cache and its registers spill):
//  1. push: this lane's candidates (mask bits 0-15 row A, 16-31 row A + 1; per row bits 0-7 bright, 8-15 dark pixels
//     x0 .. x0+7) go to the warp's queue in lane order;
//  2. drain: full batches of 32 candidates are scored, one per lane;
//  3. rolling NMS: the row pair three pairs behind, (A - 5, A - 4), is suppressed and emitted -- it needs the rows up to
//     A - 3 scored, i.e. the queue head past the tail recorded two calls ago (s_tail); in the rare case it is not (fewer
//     than 32 candidates in several rows) the partial batch is scored now.  `final`: everything is scored and the last
//     three pairs are emitted.
// Returns the new (head, tail) packed in 64 bits.
struct NmsOut {
  uint32_t* row_kp;  // list of (row 0, this strip); a row's lists are row_stride entries apart
  int* row_count;
  int row_stride;    // strips_cap
  int strip;
};
__device__ __forceinline__ void nms_emit_pair(const uint8_t* s_score, int t, int nms, unsigned emit_mask, int col0, int row_end,
                                              const NmsOut& o) {
  // rows t and t + 1 of the strip: 3x3 strict maximum over the score ring with byte-SIMD max / compare (8 pixels per lane),
  // then ordered emission into the (row, strip) keypoint lists
  const int lane = threadIdx.x & 31;
  const uint2 r0 = *reinterpret_cast<const uint2*>(s_score + ((t - 1) & (RING - 1)) * 256 + lane * 8);
  const uint2 r1 = *reinterpret_cast<const uint2*>(s_score + (t & (RING - 1)) * 256 + lane * 8);
  const uint2 r2 = *reinterpret_cast<const uint2*>(s_score + ((t + 1) & (RING - 1)) * 256 + lane * 8);
  const uint2 r3 = *reinterpret_cast<const uint2*>(s_score + ((t + 2) & (RING - 1)) * 256 + lane * 8);
  // 16-bit SIMD (VIMNMX3.U16x2 is native, byte-wide max / compare are ~7-instruction emulations): every row is split into
  // pixel pairs E0 = (p0, p2), O0 = (p1, p3), E1 = (p4, p6), O1 = (p5, p7), one score per 16-bit half
  unsigned rw[4][4];
  {
    const uint2 rr[4] = {r0, r1, r2, r3};
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      rw[q][0] = rr[q].x & M16;
      rw[q][1] = __byte_perm(rr[q].x, 0u, 0x4341);
      rw[q][2] = rr[q].y & M16;
      rw[q][3] = __byte_perm(rr[q].y, 0u, 0x4341);
    }
  }
  unsigned bits[2];
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    unsigned t[4];
    if (nms) {
      // response s - 1 strictly greater than every neighbour's (0 for non-corners): s > max(neighbours, 1)
      unsigned v[4], a[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        v[k] = __vimax3_u16x2(rw[h][k], rw[h + 2][k], 0x00010001u);  // rows above / below, same column (and the floor 1)
        a[k] = __vmaxu2(v[k], rw[h + 1][k]);                         // all three rows
      }
      unsigned prev = __shfl_up_sync(FULL, a[3], 1), next = __shfl_down_sync(FULL, a[0], 1);
      if (lane == 0) prev = 0u;   // strip column -1 / 256: never next to an emitted column
      if (lane == 31) next = 0u;
      // left / right neighbour columns of the 3-row maxima, per pixel pair
      const unsigned l_e0 = __byte_perm(prev, a[1], 0x5432) /*(p-1, p1)*/, r_e0 = a[1] /*(p1, p3)*/;
      const unsigned l_o0 = a[0] /*(p0, p2)*/, r_o0 = __byte_perm(a[0], a[2], 0x5432) /*(p2, p4)*/;
      const unsigned l_e1 = __byte_perm(a[1], a[3], 0x5432) /*(p3, p5)*/, r_e1 = a[3] /*(p5, p7)*/;
      const unsigned l_o1 = a[2] /*(p4, p6)*/, r_o1 = __byte_perm(a[2], next, 0x5432) /*(p6, p8)*/;
      // bit 15 of a half: neighbour maximum >= centre (halves stay in 0x7f01 .. 0x80ff: no borrow across)
      t[0] = __vimax3_u16x2(l_e0, r_e0, v[0]) + 0x80008000u - rw[h + 1][0];
      t[1] = __vimax3_u16x2(l_o0, r_o0, v[1]) + 0x80008000u - rw[h + 1][1];
      t[2] = __vimax3_u16x2(l_e1, r_e1, v[2]) + 0x80008000u - rw[h + 1][2];
      t[3] = __vimax3_u16x2(l_o1, r_o1, v[3]) + 0x80008000u - rw[h + 1][3];
    } else {
#pragma unroll
      for (int k = 0; k < 4; ++k) t[k] = 0x80008000u - rw[h + 1][k];  // bit 15: not a corner
    }
    // bit 15 -> pixel (0|1|4|5), bit 31 -> pixel (2|3|6|7)
    const unsigned w = ((t[0] >> 15) & 0x00010001u) | ((t[1] >> 14) & 0x00020002u) | ((t[2] >> 11) & 0x00100010u) |
                       ((t[3] >> 10) & 0x00200020u);
    bits[h] = ~(w | (w >> 14)) & emit_mask;  // emit_mask has bits 0-7 only
  }
  // one prefix scan for both rows (counts packed in 16-bit halves)
  const int cnt = __popc(bits[0]) | (__popc(bits[1]) << 16);
  int incl = cnt;
#pragma unroll
  for (int o2 = 1; o2 < 32; o2 <<= 1) {
    const int tt = __shfl_up_sync(FULL, incl, o2);
    if (lane >= o2) incl += tt;
  }
  const int excl = incl - cnt;
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int row = t + h;
    if (row >= row_end) break;
    const uint2 c = h ? r2 : r1;
    uint32_t* out = o.row_kp + (size_t) row * o.row_stride * STRIP_LIST;
    int pos = (excl >> (16 * h)) & 0xffff;
    unsigned b = bits[h];
    while (b) {
      const int j = __ffs(b) - 1;
      b &= b - 1;
      const unsigned sv = ((j < 4 ? c.x : c.y) >> (8 * (j & 3))) & 0xffu;
      out[pos++] = ((unsigned) (col0 + j) << 8) | (nms ? sv : 1u);  // without NMS the response is 0
    }
    if (lane == 31) o.row_count[(size_t) row * o.row_stride] = (incl >> (16 * h)) & 0xffff;
  }
}

__device__ __noinline__ unsigned long long push_drain_nms(const uint8_t* img, uint8_t* s_score, unsigned short* s_queue,
                                                          unsigned* s_tail, int stride, int thr, int nms, unsigned mask, int row_a,
                                                          int by, int row_end, NmsOut out, unsigned q_head, unsigned q_tail,
                                                          bool final) {
  const int lane = threadIdx.x & 31;
  // pending candidates belong to rows row_a - 8 .. row_a + 1: the ring row identifies the image row
  ScoreCtx c{img, s_score, stride, thr, row_a - 8};
  const unsigned ebase = (unsigned) (row_a & (RING - 1)) << 10;
  unsigned rest = 0u;
  bool more;
  do {
    int cnt = __popc(mask);
    int incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(FULL, incl, o);
      if (lane >= o) incl += t;
    }
    int total = __shfl_sync(FULL, incl, 31);
    if (total > QCAP - 32) {  // cannot happen on natural images (a pixel passing both polarity pre-tests): one row at a time
      rest = mask & 0xffff0000u;
      mask &= 0xffffu;
      cnt = __popc(mask);
      incl = cnt;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(FULL, incl, o);
        if (lane >= o) incl += t;
      }
      total = __shfl_sync(FULL, incl, 31);
    }
    if (total) {
      unsigned pos = q_tail + incl - cnt;
      const unsigned eb = ebase + lane * 32u;
      while (mask) {
        const unsigned i = __ffs(mask) - 1;
        mask &= mask - 1;
        s_queue[pos & (QCAP - 1)] = (unsigned short) (eb + i);
        ++pos;
      }
      q_tail += total;
      __syncwarp();
    }
    while (q_tail - q_head >= 32) {
      score_entry(c, s_queue[(q_head + lane) & (QCAP - 1)]);
      q_head += 32;
    }
    more = __any_sync(FULL, rest != 0u);
    mask = rest;
    rest = 0u;
  } while (more);
  // ---- rolling NMS ----
  const unsigned need = final ? q_tail : s_tail[1];  // tail after the pair (row_a - 4, row_a - 3) was pushed
  if ((int) (need - q_head) > 0) {                   // rows the NMS is about to read still have unscored candidates
    if ((unsigned) lane < q_tail - q_head) score_entry(c, s_queue[(q_head + lane) & (QCAP - 1)]);
    q_head = q_tail;
  }
  __syncwarp();
  if (lane == 0) {
    s_tail[1] = s_tail[0];
    s_tail[0] = q_tail;
  }
  // emit range of this strip: columns 2 .. 253 (strip 0: from 0); the overlap columns belong to the neighbour strips
  unsigned emit_mask = 0;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int k = lane * 8 + j;
    if ((k >= 2 || out.strip == 0) && k < STRIP_STRIDE + 2) emit_mask |= 1u << j;
  }
  const int col0 = out.strip * STRIP_STRIDE + lane * 8;
  const int n_pairs = final ? 3 : 1;
  for (int p = 0; p < n_pairs; ++p) {
    const int t = row_a - 5 + 2 * p;
    if (t >= by && t < row_end) {
      nms_emit_pair(s_score, t, nms, emit_mask, col0, row_end, out);
      // rows t - 1 and t are not needed any more: their ring slots are reused RING rows later
      *reinterpret_cast<uint2*>(s_score + ((t - 1) & (RING - 1)) * 256 + lane * 8) = make_uint2(0u, 0u);
      *reinterpret_cast<uint2*>(s_score + (t & (RING - 1)) * 256 + lane * 8) = make_uint2(0u, 0u);
    }
    __syncwarp();
  }
  return ((unsigned long long) q_tail << 32) | q_head;
}

// 96 registers: 4 CTAs of 5 warps per SM at KITTI width.  Shared memory is 6.1 KB per warp (score ring, candidate queue).
__global__ void __maxnreg__(96)
fast_blur_rows_kernel(const K1Args a) {
  extern __shared__ __align__(16) uint8_t smem[];
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = blockIdx.z * (blockDim.x >> 5) + (tid >> 5);  // strip index (the strips of a band are independent)
  const int rows = a.rows, cols = a.cols, stride = a.stride, thr = a.thr, BH = a.bh;  // BH is even
  if (warp >= a.n_strips) return;
  uint8_t* s_score = smem + (size_t) (tid >> 5) * WARP_SMEM;                             // [RING][256]
  unsigned short* s_queue = reinterpret_cast<unsigned short*>(s_score + RING * 256);     // [QCAP]
  unsigned* s_tail = reinterpret_cast<unsigned*>(s_score + RING * 256 + QCAP * 2);        // queue tails after the last two pushes
  const int by = blockIdx.x * BH, image = blockIdx.y;
  const int band_end = min(by + BH, rows);
  const uint8_t* img = a.images + (size_t) image * a.image_pitch;
  uint8_t* blur = a.blur + (size_t) image * a.map_slot;
  const int xs = warp * STRIP_STRIDE;     // first column of the strip
  const int x0 = xs + lane * 8;
  const bool has_left = warp > 0;  // lane 0 of strip 0 has no left halo (columns < 0)

  for (int i = lane; i < RING * 256 / 4; i += 32) reinterpret_cast<unsigned*>(s_score)[i] = 0u;
  if (lane < RING) s_tail[lane] = 0u;
  __syncwarp();

  // FAST centres: 3 <= x <= cols - 4
  unsigned colmask = 0;
#pragma unroll
  for (int j = 0; j < 8; ++j)
    if (x0 + j >= 3 && x0 + j <= cols - 4) colmask |= 1u << j;
  colmask |= colmask << 8;  // bright bits 0-7, dark bits 8-15
  const unsigned t1 = (unsigned) (thr + 1) * 0x00010001u;
  const bool store_lo = x0 < a.map_pitch, store_hi = x0 + 4 < a.map_pitch;
  unsigned q_head = 0, q_tail = 0;
  NmsOut out;
  out.row_stride = a.strips_cap;
  out.strip = warp;
  out.row_kp = a.row_kp + ((size_t) image * a.max_rows * a.strips_cap + warp) * STRIP_LIST;
  out.row_count = a.row_count + (size_t) image * a.max_rows * a.strips_cap + warp;

  // Register windows.  Raw rows (8: w0 .. w7 = image rows r - 6 .. r + 1): the compass pre-test of centre row c reads the
  // own pixels of rows c - 3 / c + 3 and the full 16-pixel window of row c, so the three oldest rows keep their own
  // pixels only.  Horizontally filtered rows: four ROW PAIRS P0 .. P3, word x = h(first row, x) | h(second row, x) << 16.
  RowOwn w0, w1, w2;
  RowRegs w3, w4, w5;
  w0.a0 = w0.a1 = w1.a0 = w1.a1 = w2.a0 = w2.a1 = 0u;
  w3.hl = w3.a0 = w3.a1 = w3.hr = 0u;
  w4 = w3;
  w5 = w3;
  unsigned P0[8], P1[8], P2[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) P0[k] = P1[k] = P2[k] = 0u;

  auto row_ptr = [&](int r) {
    const int yy = r < 0 ? 0 : (r >= rows ? rows - 1 : r);
    return img + (size_t) yy * stride;
  };
  const LaneOffs lo = lane_offs(lane, has_left);
  // software pipeline: the loads of the next row pair are issued one iteration ahead
  RawQ nxa = loadq_issue(row_ptr(by - 4), xs, lo, cols);
  RawQ nxb = loadq_issue(row_ptr(by - 3), xs, lo, cols);

  // One iteration per ROW PAIR: image rows r, r + 1 enter the windows, the rows c0 = r - 3 and c1 = r - 2 get their blur
  // output and their pre-test -- (c0, c1) = (by - 1 + 2j, by + 2j) is one (A, A + 1) pair of push_drain_nms.  The windows
  // are rotated with register moves once per pair (half of what a row-at-a-time march pays; unrolling over the window
  // period instead overflows the instruction cache).
  const int n_iter = BH / 2 + 4;
#pragma unroll 1
  for (int it = 0; it < n_iter; ++it) {
    const int r = by - 4 + 2 * it;
    const unsigned da = (unsigned) (uintptr_t) (row_ptr(r) + xs) & 7u;      // uniform over the warp
    const unsigned db = (unsigned) (uintptr_t) (row_ptr(r + 1) + xs) & 7u;
    const RowRegs w6 = assemble_row(nxa, da, lane);
    const RowRegs w7 = assemble_row(nxb, db, lane);
    if (it + 1 < n_iter) {
      nxa = loadq_issue(row_ptr(r + 2), xs, lo, cols);
      nxb = loadq_issue(row_ptr(r + 3), xs, lo, cols);
    }
    unsigned P3[8];
    {
      unsigned ha[8], hb[8];
      hblur8(w6, ha);
      hblur8(w7, hb);
#pragma unroll
      for (int k = 0; k < 8; ++k) P3[k] = __byte_perm(ha[k], hb[k], 0x5410);
    }
    const int c0 = r - 3, c1 = r - 2;
    if (c1 >= by && c0 < band_end) {
      unsigned f[8], g[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) vblur_pairs(P0[k], P1[k], P2[k], P3[k], f[k], g[k]);
      // x0 is a multiple of 4 (strips start every 252 pixels): two 4-byte stores per row
      if (c0 >= by) {
        unsigned* bo = reinterpret_cast<unsigned*>(blur + (size_t) c0 * a.map_pitch + x0);
        if (store_lo) bo[0] = pack_b2(f[0], f[1], f[2], f[3]);
        if (store_hi) bo[1] = pack_b2(f[4], f[5], f[6], f[7]);
      }
      if (c1 < band_end) {
        unsigned* bo = reinterpret_cast<unsigned*>(blur + (size_t) c1 * a.map_pitch + x0);
        if (store_lo) bo[0] = pack_b2(g[0], g[1], g[2], g[3]);
        if (store_hi) bo[1] = pack_b2(g[4], g[5], g[6], g[7]);
      }
    }
    if (c0 >= by - 1) {  // rows by - 1 .. by + BH: BH / 2 + 1 pairs
      unsigned m0 = 0u, m1 = 0u;
      if (c0 >= 3 && c0 < rows - 3) {
        unsigned mb, md;
        const RowOwn s6{w6.a0, w6.a1};
        pretest8(w0, w3, s6, t1, mb, md);
        m0 = (mb | (md << 8)) & colmask;
      }
      if (c1 >= 3 && c1 < rows - 3) {
        unsigned mb, md;
        const RowOwn s7{w7.a0, w7.a1};
        pretest8(w1, w4, s7, t1, mb, md);
        m1 = (mb | (md << 8)) & colmask;
      }
      const unsigned long long q = push_drain_nms(img + xs, s_score, s_queue, s_tail, stride, thr, a.nms, m0 | (m1 << 16), c0, by,
                                                  band_end, out, q_head, q_tail, it + 1 == n_iter);
      q_head = (unsigned) q;
      q_tail = (unsigned) (q >> 32);
    }
    // rotate the windows by one row pair
    w0 = w2;
    w1 = RowOwn{w3.a0, w3.a1};
    w2 = RowOwn{w4.a0, w4.a1};
    w3 = w5;
    w4 = w6;
    w5 = w7;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      P0[k] = P1[k];
      P1[k] = P2[k];
      P2[k] = P3[k];
    }
  }
}

// ---- exact reflect-101 blur on the 3-pixel image frame (pslam_blur7 only) ----------------------------------
__global__ void blur_border_kernel(const uint8_t* __restrict__ img, int rows, int cols, int stride,
                                   uint8_t* __restrict__ blur, int map_pitch) {
  const int n_frame = 2 * 3 * cols + 2 * 3 * (rows - 6 > 0 ? rows - 6 : 0);
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_frame) return;
  int x, y;
  if (i < 3 * cols) {
    y = i / cols;
    x = i - y * cols;
  } else if (i < 6 * cols) {
    const int k = i - 3 * cols;
    y = rows - 3 + k / cols;
    x = k % cols;
  } else {
    const int k = i - 6 * cols, rr = rows - 6;
    const int side = k / (3 * rr), kk = k % (3 * rr);
    y = 3 + kk / 3;
    x = side == 0 ? kk % 3 : cols - 3 + kk % 3;
  }
  if (y < 0 || y >= rows) return;
  const int kq[7] = {18, 34, 49, 55, 49, 34, 18};
  int v = 0;
  for (int dy = -3; dy <= 3; ++dy) {
    const int yy = reflect101(y + dy, rows);
    int h = 0;
    for (int dx = -3; dx <= 3; ++dx) h += kq[dx + 3] * img[(size_t) yy * stride + reflect101(x + dx, cols)];
    v += kq[dy + 3] * h;
  }
  blur[(size_t) y * map_pitch + x] = (uint8_t) min(255, (v + 32768) >> 16);
}

// -------------------------------------------------------------------------------------------------------------
// K2: gather + select.  grid (regions, images), one CTA of SEL_THREADS threads.
// -------------------------------------------------------------------------------------------------------------
constexpr int SEL_THREADS = 128;      // throughput launches (many images).  Measured us / image: 64 threads 0.373, 96 0.338, 128 0.316, 192 0.414,
                                      // 256 0.450 (the gather is one thread per region row; 12 kB of sort scratch per CTA: 16 CTAs = 64 warps per SM)
constexpr int SEL_THREADS_WIDE = 256;  // latency launches (one or two images): the gather walks all rows of a region at once

struct RespGreater {
  __device__ __forceinline__ bool operator()(const uint32_t& x, const uint32_t& y) const {
    return (x & 0xffu) > (y & 0xffu);
  }
};

template <int NT>
__global__ void __launch_bounds__(NT)
bin_select_kernel(const int* __restrict__ row_count, const uint32_t* __restrict__ row_kp, int strips_cap, int max_rows,
                  const uint8_t* __restrict__ mask, int mask_pitch, int mask_invert, int rows, int cols, int nh, int nv,
                  float pixel_rows_per_detector, float pixel_cols_per_detector, uint32_t* __restrict__ raw,
                  int max_raw_per_bin, int max_bins, int* __restrict__ raw_count, int* __restrict__ sel_count,
                  unsigned long long quota, int sort_cap, int* __restrict__ flags, const int* __restrict__ bounds) {
  extern __shared__ __align__(16) uint32_t s_sort[];
  __shared__ int s_warp[33];
  __shared__ int s_rbegin;
  const int tid = threadIdx.x;
  const int bin = blockIdx.x, image = blockIdx.y;
  uint32_t* seg = raw + ((size_t) image * max_bins + bin) * max_raw_per_bin;
  // pixel range of this region: the (r, c) -> region rule of binned.cpp:83-90 is the same for every image of a launch;
  // the launcher evaluates it once (same fp32 arithmetic) instead of every CTA walking all rows and columns
  const int4 bd = __ldg(reinterpret_cast<const int4*>(bounds) + bin);
  const int rbegin = bd.x, rend = bd.y, cbegin = bd.z, cend = bd.w;
  // ordered gather: one thread per image row, rows in chunks of NT.  A row's keypoints come as one list per
  // 252-pixel strip (K1), strips in column order: only the strips that overlap the region's columns are walked.
  const int s_first = cbegin < STRIP_STRIDE + 2 ? 0 : (cbegin - 2) / STRIP_STRIDE;
  const int s_last = cend - 1 < STRIP_STRIDE + 2 ? 0 : (cend - 1 - 2) / STRIP_STRIDE;
  int running = 0;
  for (int base = rbegin; base < rend; base += NT) {
    const int r = base + tid;
    int cnt = 0;
    if (r < rend) {
      for (int st = s_first; st <= s_last; ++st) {
        const size_t o = ((size_t) image * max_rows + r) * strips_cap + st;
        const int n_r = row_count[o];
        const uint32_t* src = row_kp + o * STRIP_LIST;
        // four list entries per load (lists are 16-byte aligned): the walk is a chain of dependent L2 accesses
        bool done = false;
        for (int k0 = 0; k0 < n_r && !done; k0 += 4) {
          const uint4 v = __ldg(reinterpret_cast<const uint4*>(src + k0));
          const uint32_t e4[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int k = k0 + u;
            if (done || k >= n_r) {
              done = true;
              continue;
            }
            const int c = (int) (e4[u] >> 8);
            if (c < cbegin) continue;
            if (c >= cend) {
              done = true;
              continue;
            }
            if (mask && ((mask[(size_t) r * mask_pitch + c] == 0) != (mask_invert != 0))) continue;
            ++cnt;
          }
        }
      }
    }
    int total;
    const int off = block_exclusive_scan<NT>(cnt, s_warp, &total);
    if (cnt) {
      int o = running + off;
      for (int st = s_first; st <= s_last; ++st) {
        const size_t ol = ((size_t) image * max_rows + r) * strips_cap + st;
        const int n_r = row_count[ol];
        const uint32_t* src = row_kp + ol * STRIP_LIST;
        bool done = false;
        for (int k0 = 0; k0 < n_r && !done; k0 += 4) {
          const uint4 v = __ldg(reinterpret_cast<const uint4*>(src + k0));
          const uint32_t e4[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int k = k0 + u;
            if (done || k >= n_r) {
              done = true;
              continue;
            }
            const uint32_t e = e4[u];
            const int c = (int) (e >> 8);
            if (c < cbegin) continue;
            if (c >= cend) {
              done = true;
              continue;
            }
            if (mask && ((mask[(size_t) r * mask_pitch + c] == 0) != (mask_invert != 0))) continue;
            if (o < max_raw_per_bin) seg[o] = ((uint32_t) (r * cols + c) << 8) | (e & 0xffu);
            ++o;
          }
        }
      }
    }
    running += total;
  }
  __syncthreads();
  int n = running;
  if (n > max_raw_per_bin) {
    if (tid == 0) atomicOr(flags, PSLAM_FLAG_RAW_OVERFLOW);
    n = max_raw_per_bin;
  }
  // selection (binned.cpp:180-200): keep all when fewer than the quota, else unstable std::sort by response and keep
  // the first `quota` -- replayed move for move (libstdcxx_sort.h), warp-parallel when the region fits shared memory
  int kept = n;
  if ((unsigned long long) n >= quota) {
    kept = (int) quota;
    if (kept > 0) {
      if (n <= sort_cap) {
        unsigned short* s_rpos = reinterpret_cast<unsigned short*>(s_sort + sort_cap);
        for (int i = tid; i < n; i += NT) s_sort[i] = seg[i];
        __syncthreads();
        // the partitions of a recursion level of the replay run on all warps of the CTA (at most sort_cap / 17 live segments)
        constexpr int SEL_SEG_CAP = 128;
        __shared__ int s_seg[2 * 3 * SEL_SEG_CAP], s_cnt[3];
        const int se = pslam_sort::block_std_sort_prefix<NT, SEL_SEG_CAP>(s_sort, s_rpos, n, kept, RespGreater(), s_seg, s_cnt);
        if (tid == 0) s_rbegin = se;
        __syncthreads();
        pslam_sort::block_final_positions<NT>(s_sort, s_rbegin, kept, seg, RespGreater());
      } else if (tid == 0) {
        pslam_sort::std_sort_prefix(seg, n, kept, RespGreater());
      }
    }
  }
  if (tid == 0) {
    raw_count[image * max_bins + bin] = n;
    sel_count[image * max_bins + bin] = kept;
  }
}

}  // namespace

// ---- host-side launchers -----------------------------------------------------------------------------------
int pslam_k_strips_cap(int max_cols) { return k1_strips(max_cols); }

int pslam_k_fast_blur(pslam_ctx* ctx, const uint8_t* d_images, long long image_pitch, int n_images,
                      int rows, int cols, int stride, int thr, int nms) {
  const int n_strips = k1_strips(cols);
  if (n_strips > 16 || n_strips > ctx->strips_cap)
    return pslam_set_error(ctx, PSLAM_E_CAPACITY, "images wider than 4032 pixels (or than pslam_limits.max_cols) are not supported", cudaSuccess);
  thr = thr < 0 ? 0 : (thr > 255 ? 255 : thr);
  K1Args a;
  a.images = d_images;
  a.image_pitch = image_pitch;
  a.rows = rows;
  a.cols = cols;
  a.stride = stride;
  a.thr = thr;
  a.nms = nms;
  a.blur = ctx->d_blur;
  a.map_pitch = ctx->map_pitch;
  a.map_slot = (long long) ctx->map_slot;
  a.row_count = ctx->d_row_count;
  a.row_kp = ctx->d_row_kp;
  a.strips_cap = ctx->strips_cap;
  a.max_rows = ctx->lim.max_rows;
  a.n_strips = n_strips;
  // band height: every band re-marches 8 halo rows, so taller bands waste less; the grid still has to fill the GPU
  // (4 CTAs per SM resident) several times over for the tail to stay small.  Throughput mode: ~48-row bands
  // (PSLAM_K1_BH overrides, for tuning runs).
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ctx->device);
  static const int bh_env = [] {
    const char* e = getenv("PSLAM_K1_BH");
    return e ? atoi(e) : 0;
  }();
  static const int wpc_env = [] {
    const char* e = getenv("PSLAM_K1_WPC");
    return e ? atoi(e) : 0;
  }();
  const int bh_target = bh_env > 0 ? bh_env : 94;
  int n_bands = (rows + bh_target - 1) / bh_target;
  a.bh = ((rows + n_bands - 1) / n_bands + 1) & ~1;  // even: the rows by - 1 .. by + bh are processed in pairs
  // latency mode: a launch over one or two images (the per-frame adaptor) would occupy a fraction of the SMs with long
  // serial marches; shorter bands (>= 8 rows) put about one CTA on every SM instead
  if ((long long) n_images * n_bands < sms) {
    const int want = (sms + n_images - 1) / n_images;
    int bh = ((rows + want - 1) / want + 1) & ~1;
    if (bh < 8) bh = 8;
    if (bh < a.bh) {
      a.bh = bh;
      n_bands = (rows + bh - 1) / bh;
    }
  }
  // the strips of a band never synchronise: warps per CTA is a scheduling choice.  Measured at KITTI width (5 strips):
  // 5 warps per CTA 2.04, 1 warp per CTA 2.15, 2 warps per CTA 2.49 us / image
  int wpc = wpc_env > 0 ? wpc_env : n_strips;
  if (wpc > n_strips) wpc = n_strips;
  const size_t smem = (size_t) wpc * WARP_SMEM;
  // the limit is a property of the FUNCTION (per device), shared by every context of the process: it is only ever raised --
  // a per-context cache would let one context lower it under another one's launches ("invalid argument")
  static size_t k1_smem_limit[64] = {0};
  static std::mutex k1_smem_mutex;  // contexts of different host threads may launch on the same device
  const int dev_slot = ctx->device & 63;
  if (smem > 48 * 1024) {
    std::lock_guard<std::mutex> lock(k1_smem_mutex);
    if (smem > k1_smem_limit[dev_slot]) {
      PSLAM_CUDA_TRY(ctx, cudaFuncSetAttribute(fast_blur_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
      k1_smem_limit[dev_slot] = smem;
    }
  }
  dim3 grid(n_bands, n_images, (n_strips + wpc - 1) / wpc);
  fast_blur_rows_kernel<<<grid, 32 * wpc, smem, ctx->stream>>>(a);
  PSLAM_LAUNCH_CHECK(ctx, "fast_blur_rows_kernel");
  return PSLAM_OK;
}

int pslam_k_blur_border(pslam_ctx* ctx, const uint8_t* d_image, int rows, int cols, int stride) {
  const int n_frame = 6 * cols + 6 * (rows > 6 ? rows - 6 : 0);
  blur_border_kernel<<<(n_frame + 255) / 256, 256, 0, ctx->stream>>>(d_image, rows, cols, stride, ctx->d_blur, ctx->map_pitch);
  PSLAM_LAUNCH_CHECK(ctx, "blur_border_kernel");
  return PSLAM_OK;
}

int pslam_k_bin_select(pslam_ctx* ctx, int n_images, int rows, int cols, int nh, int nv,
                       unsigned long long quota, const uint8_t* d_mask, int mask_invert) {
  // float arithmetic of IntensityFeatureExtractorBinned_::init (binned.cpp:49-52)
  const float pr = static_cast<float>(rows) / static_cast<float>((size_t) nv);
  const float pc = static_cast<float>(cols) / static_cast<float>((size_t) nh);
  // region bounds (binned.cpp:83-90): rows with floor(r / pr) == rb; columns with (size_t)((float) row_region + c / pc) ==
  // bin, monotone in c.  Same fp32 operations as the reference's LUT, evaluated once per image geometry.
  if (ctx->sel_bounds_key[0] != rows || ctx->sel_bounds_key[1] != cols || ctx->sel_bounds_key[2] != nh || ctx->sel_bounds_key[3] != nv) {
    std::vector<int> bd((size_t) 4 * nh * nv);
    for (int bin = 0; bin < nh * nv; ++bin) {
      const int rb = bin / nh;
      int rbegin = rows, rend = 0, cbegin = cols, cend = 0;
      for (int r = 0; r < rows; ++r)
        if ((int) floorf((float) r / pr) == rb) {
          rbegin = r < rbegin ? r : rbegin;
          rend = r + 1;
        }
      volatile float row_region = (float) rb * (float) nh;
      for (int c = 0; c < cols; ++c) {
        volatile float q = (float) c / pc;
        volatile float f = (float) (unsigned) row_region + q;
        if ((unsigned) f == (unsigned) bin) {
          cbegin = c < cbegin ? c : cbegin;
          cend = c + 1;
        }
      }
      bd[4 * bin] = rbegin;
      bd[4 * bin + 1] = rend;
      bd[4 * bin + 2] = cbegin;
      bd[4 * bin + 3] = cend;
    }
    PSLAM_CUDA_TRY(ctx, cudaMemcpyAsync(ctx->d_sel_bounds, bd.data(), bd.size() * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    PSLAM_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));  // bd is a local
    ctx->sel_bounds_key[0] = rows;
    ctx->sel_bounds_key[1] = cols;
    ctx->sel_bounds_key[2] = nh;
    ctx->sel_bounds_key[3] = nv;
  }
  dim3 grid(nh * nv, n_images);
  const int sort_cap = ctx->lim.max_raw_per_bin < 2048 ? ctx->lim.max_raw_per_bin : 2048;
  // one or two images (the per-frame adaptor): 9-18 CTAs cannot fill the GPU, so each gets four times the threads
  const bool wide = (long long) nh * nv * n_images <= 64;
  auto kernel = wide ? bin_select_kernel<SEL_THREADS_WIDE> : bin_select_kernel<SEL_THREADS>;
  kernel<<<grid, wide ? SEL_THREADS_WIDE : SEL_THREADS, (size_t) sort_cap * 6, ctx->stream>>>(
    ctx->d_row_count, ctx->d_row_kp, ctx->strips_cap, ctx->lim.max_rows, d_mask, ctx->map_pitch, mask_invert, rows, cols, nh, nv, pr,
    pc, ctx->d_raw, ctx->lim.max_raw_per_bin, ctx->lim.max_bins, ctx->d_raw_count, ctx->d_sel_count, quota, sort_cap,
    ctx->d_flags, ctx->d_sel_bounds);
  PSLAM_LAUNCH_CHECK(ctx, "bin_select_kernel");
  return PSLAM_OK;
}

/* Native C99 caller of the drop-in boundary (include/pslam_cuda.h): no Python, no C++, no torch -- what a binding inside the
 * reference's build would look like from the outside.  Built and run by tests/test_cpu_abi.py (without a GPU: pslam_create must
 * refuse with PSLAM_E_CUDA, there is no CPU fallback) and tests/test_gpu_batch.py (with one: extraction and the stereo adaptor on
 * a synthetic pair, deterministic across calls).
 *   gcc -std=c99 -Iinclude tests/native/abi_smoke.c -Lsrrg2_proslam_b200 -lpslam_cuda -Wl,-rpath,$PWD/srrg2_proslam_b200 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "pslam_cuda.h"

#define ROWS 240
#define COLS 416

static unsigned lcg(unsigned* s) { return *s = *s * 1664525u + 1013904223u; }

/* blocky texture with noise; the right image is the left one shifted by `disp` pixels */
static void make_pair(unsigned char* left, unsigned char* right, int disp) {
  static unsigned char tex[ROWS][COLS + 64];
  unsigned s = 12345u;
  int r, c;
  for (r = 0; r < ROWS; r += 8)
    for (c = 0; c < COLS + 64; c += 8) {
      const unsigned char v = (unsigned char) (40 + lcg(&s) % 160);
      int y, x;
      for (y = r; y < r + 8 && y < ROWS; ++y)
        for (x = c; x < c + 8 && x < COLS + 64; ++x) tex[y][x] = (unsigned char) (v + lcg(&s) % 9);
    }
  for (r = 0; r < ROWS; ++r)
    for (c = 0; c < COLS; ++c) {
      left[r * COLS + c] = tex[r][c + disp];
      right[r * COLS + c] = tex[r][c + 2 * disp];
    }
}

int main(void) {
  pslam_limits lim;
  pslam_ctx* ctx = NULL;
  int rc;
  memset(&lim, 0, sizeof lim);
  lim.max_images = 2;
  lim.max_rows = ROWS;
  lim.max_cols = COLS;
  lim.max_features = 2048;
  lim.max_raw_per_bin = 8192;
  lim.max_bins = 9;
  printf("%s\n", pslam_version());
  rc = pslam_create(0, &lim, &ctx);
  if (rc == PSLAM_E_CUDA) {
    printf("NO_DEVICE: %s\n", ctx ? pslam_last_error(ctx) : "pslam_create refused");
    pslam_destroy(ctx);
    return 3; /* the CPU test expects exactly this */
  }
  if (rc != PSLAM_OK) {
    printf("pslam_create failed: %d\n", rc);
    return 1;
  }
  {
    unsigned char* L = (unsigned char*) malloc(ROWS * COLS);
    unsigned char* R = (unsigned char*) malloc(ROWS * COLS);
    float *xy = (float*) malloc(sizeof(float) * 2 * 2048), *resp = (float*) malloc(sizeof(float) * 2048);
    float* inten = (float*) malloc(sizeof(float) * 2048);
    unsigned char *desc = (unsigned char*) malloc(32 * 2048), *desc2 = (unsigned char*) malloc(32 * 2048);
    float *uvuv = (float*) malloc(sizeof(float) * 4 * 2048), *uvuv2 = (float*) malloc(sizeof(float) * 4 * 2048);
    pslam_extract_cfg e = {15.0f, 1, 600, 3, 3};
    pslam_match_cfg m;
    int n1, n2, s1, s2, i, ok = 1;
    memset(&m, 0, sizeof m);
    m.maximum_descriptor_distance = 100.0f;
    m.maximum_distance_ratio_to_second_best = 0.8f;
    m.maximum_disparity_pixels = 100;
    m.epipolar_line_thickness_pixels = 0;
    make_pair(L, R, 12);
    n1 = pslam_extract_binned(ctx, L, ROWS, COLS, COLS, &e, NULL, 2048, xy, resp, inten, desc);
    n2 = pslam_extract_binned(ctx, L, ROWS, COLS, COLS, &e, NULL, 2048, xy, resp, inten, desc2);
    if (n1 <= 50 || n1 != n2 || memcmp(desc, desc2, 32 * (size_t) n1) != 0) ok = 0;
    s1 = pslam_stereo_adaptor(ctx, L, R, ROWS, COLS, COLS, &e, &m, 2048, uvuv, inten, desc);
    s2 = pslam_stereo_adaptor(ctx, L, R, ROWS, COLS, COLS, &e, &m, 2048, uvuv2, inten, desc2);
    if (s1 <= 20 || s1 != s2 || memcmp(uvuv, uvuv2, sizeof(float) * 4 * (size_t) s1) != 0) ok = 0;
    for (i = 0; ok && i < s1; ++i) { /* the right image is the left one shifted by 12 px: same row, disparity 12 */
      if (uvuv[4 * i + 1] != uvuv[4 * i + 3] || uvuv[4 * i] - uvuv[4 * i + 2] != 12.0f) ok = 0;
    }
    printf("features %d, stereo points %d, kernel launches %lld, %s\n", n1, s1, pslam_launch_count(ctx), ok ? "OK" : "MISMATCH");
    if (!ok) printf("last error: %s\n", pslam_last_error(ctx));
    free(L); free(R); free(xy); free(resp); free(inten); free(desc); free(desc2); free(uvuv); free(uvuv2);
    pslam_destroy(ctx);
    return ok ? 0 : 2;
  }
}

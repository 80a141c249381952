/* oracle/dic_oracle.c -- TEST INFRASTRUCTURE (oracle), not product code.
 *
 * CPU restatement, in plain C, of the reference's CPU digital-image-correlation path
 * (namascar/correlation; citations are file:line into the reference tree):
 *
 *   pyramid build            pyramid_class.cpp:52-134
 *   per-level point lists    pyramid_class.cpp:289-323, centre :325-362, u,v scaling :260-287
 *   deformation models       model_class.cpp:48-202 (U, UV, UVQ, UVUxUyVxVy)
 *   bicubic coefficients     interpolation_class.cpp:243-336, matrix :539-558
 *   bicubic evaluation       interpolation_class.cpp:79-138 (bilinear :140-195, nearest :197-226)
 *   residual / A, b, chi     interpolation_class.cpp:671-764
 *   thread fan-out / fan-in  correlation_class.cpp:131-300
 *   LM state machine         correlation_class.cpp:349-640, solve step :642-768
 *   pixel-list builders      manager_class.cpp:1596-1614 (rect), :816-940 (annulus),
 *                            polygon_class.cpp (blob: ear clipping + scanline)
 *
 * It is pinned against the UNMODIFIED reference compiled into oracle/_ref/libdic_ref.so
 * (tests/test_oracle.py, bit-for-bit on parameters, chi, iterations, A, b, pyramid
 * levels and point lists) and against the golden fixtures generated from that build
 * (tests/golden/). Deliberate differences, none of which changes an in-bounds result:
 *   - 64-bit indices and on-the-fly bicubic coefficients instead of the reference's lazy
 *     `int`-indexed per-image cache (pyramid_class.cpp:180-187) -- the cache is a memo of a
 *     pure function of the image, so values are identical; it lets 16384^2 run.
 *     Consequence: the cache-poisoning after an out-of-image error
 *     (interpolation_class.cpp:245-250) is not reproduced; after error 2 only the error
 *     code is comparable.
 *   - optional double accumulators for A, b, chi (accum_double) to arbitrate the reference's
 *     own thread-count dependent float absorption;
 *   - fitting model 4 = 12-parameter quadratic warp, OUR extension written in the pattern of
 *     model_class.cpp:150-202 (PARITY UNPINNED: nothing in the reference covers it). Its
 *     normal equations are Jacobi-equilibrated before the QR solve (second-order columns are
 *     ~1e12 larger than the translation columns at 4096^2).
 *   - the 6x6 solve restates Eigen 3.4.0 colPivHouseholderQr (oracle/qr_colpiv.h; parity
 *     unpinned at that boundary, shared with the _ref build so the two agree).
 *   - optional fp64 solve of the same damped system (orc_set_solve_double) to arbitrate the solver the way
 *     accum_double arbitrates the accumulators: the fp32 Householder QR of the UNEQUILIBRATED normal equations
 *     (cond ~1e4..1e5 for a 125^2 subset) errs by ~cond * eps of the step, and the reported chi -- evaluated
 *     at the point that step leads to -- moves by up to 2e-4 relative with it (tools/lm_trace.py).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load this library.
 */
#define _GNU_SOURCE
#include <float.h>
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "qr_colpiv.h"

#define ORC_MAXP 12
#define ORC_MAXLEV 12

enum { IM_NEAREST = 0, IM_BILINEAR = 1, IM_BICUBIC = 2 };
enum { FM_U = 0, FM_UV = 1, FM_UVQ = 2, FM_AFFINE = 3, FM_QUAD = 4 };
enum {
  ERR_NONE = 0, ERR_MODEL_OOB = 1, ERR_INTERP_OOB = 2, ERR_MAX_ITERS = 3, ERR_BAD_DOMAIN = 4
};

typedef struct {
  float params[ORC_MAXP];
  float chi;
  int number_of_points;
  int iterations;
  int error_code;
  int error_status;
  float und_center_x;
  float und_center_y;
  double seconds;
  /* extras the reference does not report (work accounting) */
  int evaluations[ORC_MAXLEV];
  int iterations_per_level[ORC_MAXLEV];
  long points_per_level[ORC_MAXLEV];
  double pixel_evaluations;
} orc_result;

typedef struct {
  uint8_t *lev[ORC_MAXLEV];
  int rows, cols;
  int n;
} orc_pyr;

typedef struct orc_engine {
  int n_threads, real_threads, accum_double;
  int colors;       /* 1 (monochrome) or 3 (interleaved colour image), set before the images (orc_set_colors) */
  int solve_double; /* arbitration variant: the damped system solved in fp64 instead of the fp32 column-pivoted QR */
  int interp, model, np;
  float precision;
  int max_iters;
  int start, step, stop;
  orc_pyr und, def, nxt;
  float *xy[ORC_MAXLEV];
  long npts[ORC_MAXLEV];
  float cx[ORC_MAXLEV], cy[ORC_MAXLEV];
  int reached_iterations;
  float last_good_chi;
  int error_status, error_code;
  /* evaluation outputs */
  float A[ORC_MAXP * ORC_MAXP], b[ORC_MAXP], chi;
} orc_engine;

static int model_nparams(int m) {
  /* model_class.cpp:216-231 (+ our 12-parameter extension) */
  switch (m) {
  case FM_U: return 1;
  case FM_UV: return 2;
  case FM_UVQ: return 3;
  case FM_AFFINE: return 6;
  case FM_QUAD: return 12;
  }
  return -1;
}

/* ------------------------------------------------------------------ pyramid */

static void pyr_free(orc_pyr *p) {
  for (int i = 0; i < ORC_MAXLEV; ++i) { free(p->lev[i]); p->lev[i] = NULL; }
  p->rows = p->cols = p->n = 0;
}

/* pyramid_class.cpp:52-134: 5x5 kernel [.05 .25 .4 .25 .05]^2, weights formed as fp32
 * products at run time (:83-90), 25 sequential mul+add in dj-outer/di-inner order (:109-117),
 * truncation to u8 (:118-119), 1-px zero border (zero-initialised target, loops 1..n-2). */
static void pyr_build(orc_pyr *p, const uint8_t *img, int rows, int cols, int stop, int colors) {
  pyr_free(p);
  p->rows = rows; p->cols = cols; p->n = stop + 1;
  p->lev[0] = (uint8_t *)malloc((size_t)rows * cols * colors);
  memcpy(p->lev[0], img, (size_t)rows * cols * colors);
  const float km[5] = {0.05f, 0.25f, 0.4f, 0.25f, 0.05f};
  float kernel[25];
  for (int i = 0; i < 5; ++i)
    for (int j = 0; j < 5; ++j) kernel[5 * j + i] = km[i] * km[j];
  int sc = cols, sr = rows;
  for (int l = 1; l <= stop; ++l) {
    long sstep = (long)sc * colors;      /* :93 */
    int tc = sc / 2, tr = sr / 2;
    long tstep = sstep / 2;              /* :94 (== tc * colors for the even widths the colour path is used with) */
    uint8_t *dst = (uint8_t *)calloc((size_t)tc * tr * colors + colors, 1);
    const uint8_t *src = p->lev[l - 1];
    for (int tj = 1; tj < tr - 1; ++tj)
      for (int ti = 1; ti < tc - 1; ++ti)
        for (int c = 0; c < colors; ++c) {
          int si = ti * 2, sj = tj * 2;
          float addition = 0.f;
          for (int dj = -2; dj <= 2; ++dj)
            for (int di = -2; di <= 2; ++di) {
              uint8_t s = src[sstep * (sj + dj) + (long)(si + di) * colors + c];
              float ker = kernel[(2 + dj) * 5 + (2 + di)];
              addition += (float)s * ker;
            }
          dst[tstep * tj + (long)ti * colors + c] = (uint8_t)addition;
        }
    p->lev[l] = dst;
    sr = tr; sc = tc;
  }
}

/* ------------------------------------------------------------ engine set-up */

orc_engine *orc_create(int n_threads, int interp, int model, float precision, int max_iters,
                       int start, int step, int stop, int accum_double, int real_threads) {
  orc_engine *e = (orc_engine *)calloc(1, sizeof(orc_engine));
  e->n_threads = n_threads < 1 ? 1 : n_threads;
  e->real_threads = real_threads;
  e->accum_double = accum_double;
  e->interp = interp; e->model = model; e->np = model_nparams(model);
  e->precision = precision; e->max_iters = max_iters;
  e->start = start; e->step = step < 1 ? 1 : step; e->stop = stop;
  e->colors = 1;
  return e;
}

/* number_of_colors of the reference (manager_class.cpp:99-109): 3 for interleaved colour images */
void orc_set_colors(orc_engine *e, int colors) { e->colors = colors == 3 ? 3 : 1; }

static void free_points(orc_engine *e) {
  for (int i = 0; i < ORC_MAXLEV; ++i) { free(e->xy[i]); e->xy[i] = NULL; e->npts[i] = 0; }
}

void orc_destroy(orc_engine *e) {
  if (!e) return;
  pyr_free(&e->und); pyr_free(&e->def); pyr_free(&e->nxt);
  free_points(e);
  free(e);
}

void orc_set_und_image(orc_engine *e, const uint8_t *img, int rows, int cols) {
  pyr_build(&e->und, img, rows, cols, e->stop, e->colors);
}
void orc_set_def_image(orc_engine *e, const uint8_t *img, int rows, int cols) {
  pyr_build(&e->def, img, rows, cols, e->stop, e->colors);
}
void orc_set_nxt_image(orc_engine *e, const uint8_t *img, int rows, int cols) {
  pyr_build(&e->nxt, img, rows, cols, e->stop, e->colors);
}
/* pyramid_class.cpp:211-258: pointer rotation */
void orc_und_from_def(orc_engine *e) {
  pyr_free(&e->und); e->und = e->def; memset(&e->def, 0, sizeof(orc_pyr));
}
void orc_def_from_nxt(orc_engine *e) {
  pyr_free(&e->def); e->def = e->nxt; memset(&e->nxt, 0, sizeof(orc_pyr));
}

/* pyramid_class.cpp:289-323 (lists), :325-362 (centres) */
void orc_set_points(orc_engine *e, const float *xy, long n, int use_center, float cx, float cy) {
  free_points(e);
  e->xy[0] = (float *)malloc(sizeof(float) * 2 * (size_t)(n > 0 ? n : 1));
  memcpy(e->xy[0], xy, sizeof(float) * 2 * (size_t)n);
  e->npts[0] = n;
  int prev = 0;
  int first = (e->start == 0 ? e->step : e->start);
  for (int l = first; l <= e->stop; l += e->step) {
    int mag = 1 << (l - prev);
    float maginv = 1.f / (float)mag;
    long np = e->npts[prev];
    float *dst = (float *)malloc(sizeof(float) * 2 * (size_t)(np > 0 ? np : 1));
    long m = 0;
    const float *src = e->xy[prev];
    for (long i = 0; i < np; ++i) {
      int ix = (int)(src[2 * i] + 0.5f);
      int iy = (int)(src[2 * i + 1] + 0.5f);
      if (ix % mag == 0 && iy % mag == 0) {
        dst[2 * m] = src[2 * i] * maginv;
        dst[2 * m + 1] = src[2 * i + 1] * maginv;
        ++m;
      }
    }
    e->xy[l] = dst; e->npts[l] = m;
    prev = l;
  }
  if (!use_center) {
    float sx = 0.f, sy = 0.f; /* sequential fp32 sums, :332-338 */
    for (long i = 0; i < n; ++i) { sx += xy[2 * i]; sy += xy[2 * i + 1]; }
    cx = sx / (float)n; cy = sy / (float)n;
  }
  e->cx[0] = cx; e->cy[0] = cy;
  for (int l = first; l <= e->stop; l += e->step) {
    float maginv = 1.f / (float)(1 << l);
    e->cx[l] = cx * maginv; e->cy[l] = cy * maginv;
  }
}

long orc_level_num_points(orc_engine *e, int level) { return e->npts[level]; }
void orc_level_points(orc_engine *e, int level, float *out) {
  memcpy(out, e->xy[level], sizeof(float) * 2 * (size_t)e->npts[level]);
}
void orc_level_center(orc_engine *e, int level, float *cx, float *cy) {
  *cx = e->cx[level]; *cy = e->cy[level];
}
void orc_pyramid_level(orc_engine *e, int which, int level, uint8_t *out, int *rows, int *cols) {
  orc_pyr *p = which == 0 ? &e->und : (which == 1 ? &e->def : &e->nxt);
  int r = p->rows / (1 << level), c = p->cols / (1 << level);
  *rows = r; *cols = c;
  if (out) memcpy(out, p->lev[level], (size_t)r * c * e->colors);
}

/* pyramid_class.cpp:260-287: only u, v scale; the quadratic extension scales the
 * second-order terms by the inverse factor. */
static void translate_params(const orc_engine *e, float *p, int src, int dst) {
  float mag;
  if (dst - src > 0) mag = 1.f / (float)(1 << (dst - src));
  else mag = (float)(1 << (-dst + src));
  int lim = e->np < 2 ? e->np : 2;
  for (int i = 0; i < lim; ++i) p[i] *= mag;
  if (e->model == FM_QUAD) {
    float inv = 1.f / mag;
    for (int i = 6; i < 12; ++i) p[i] *= inv;
  }
}

/* ---------------------------------------------------------- bicubic machinery */

/* interpolation_class.cpp:539-558: exact inverse of the Hermite constraint matrix. */
static const float BICUBIC_M[256] = {
    16,  -20,  -20, 25,  16,   8,   -20, -10,  16,  -20, 8,   -10,  16,   8,
    8,   4,    -48, 48,  60,   -60, -32, -20,  40,  25,  -48, 48,   -24,  24,
    -32, -20,  -16, -10, 36,   -36, -45, 45,   20,  16,  -25, -20,  36,   -36,
    18,  -18,  20,  16,  10,   8,   -8,  8,    10,  -10, -4,  -4,   5,    5,
    -8,  8,    -4,  4,   -4,   -4,  -2,  -2,   -48, 60,  48,  -60,  -48,  -24,
    48,  24,   -32, 40,  -20,  25,  -32, -16,  -20, -10, 144, -144, -144, 144,
    96,  60,   -96, -60, 96,   -96, 60,  -60,  64,  40,  40,  25,   -108, 108,
    108, -108, -60, -48, 60,   48,  -72, 72,   -45, 45,  -40, -32,  -25,  -20,
    24,  -24,  -24, 24,  12,   12,  -12, -12,  16,  -16, 10,  -10,  8,    8,
    5,   5,    36,  -45, -36,  45,  36,  18,   -36, -18, 20,  -25,  16,   -20,
    20,  10,   16,  8,   -108, 108, 108, -108, -72, -45, 72,  45,   -60,  60,
    -48, 48,   -40, -25, -32,  -20, 81,  -81,  -81, 81,  45,  36,   -45,  -36,
    45,  -45,  36,  -36, 25,   20,  20,  16,   -18, 18,  18,  -18,  -9,   -9,
    9,   9,    -10, 10,  -8,   8,   -5,  -5,   -4,  -4,  -8,  10,   8,    -10,
    -8,  -4,   8,   4,   -4,   5,   -4,  5,    -4,  -2,  -4,  -2,   24,   -24,
    -24, 24,   16,  10,  -16,  -10, 12,  -12,  12,  -12, 8,   5,    8,    5,
    -18, 18,   18,  -18, -10,  -8,  10,  8,    -9,  9,   -9,  9,    -5,   -4,
    -5,  -4,   4,   -4,  -4,   4,   2,   2,    -2,  -2,  2,   -2,   2,    -2,
    1,   1,    1,   1};

const float *orc_bicubic_matrix(void) { return BICUBIC_M; }

/* interpolation_class.cpp:243-336. Column x of colour channel c lives at byte x * mult (+ add) of its row:
 * monochrome mult = 1; for colour images the reference's bicubic and bilinear coefficient builders index with
 * `color = number_of_colors + color_in; index_ix = ix * color` (:268-273, :356-359), i.e. mult = 3 + c, add = 0 --
 * correct for channel 0 only; restated as executed. Nearest uses ix * number_of_colors + color_in (:391-398). */
static void bicubic_coeffs(const uint8_t *img, long step, int x, int y, int mult, float *a) {
  const uint8_t *r0 = img + step * (y - 1), *r1 = img + step * y, *r2 = img + step * (y + 1),
                *r3 = img + step * (y + 2);
  int x0 = (x - 1) * mult, x1 = x * mult, x2 = (x + 1) * mult, x3 = (x + 2) * mult;
  float w00 = r0[x0], w01 = r1[x0], w02 = r2[x0], w03 = r3[x0];
  float w10 = r0[x1], w11 = r1[x1], w12 = r2[x1], w13 = r3[x1];
  float w20 = r0[x2], w21 = r1[x2], w22 = r2[x2], w23 = r3[x2];
  float w30 = r0[x3], w31 = r1[x3], w32 = r2[x3], w33 = r3[x3];
  float v[16];
  v[0] = w11; v[1] = w21; v[2] = w12; v[3] = w22;
  v[4] = (w21 - w01) / 2.f; v[5] = (w31 - w11) / 2.f;
  v[6] = (w22 - w02) / 2.f; v[7] = (w32 - w12) / 2.f;
  v[8] = (w12 - w10) / 2.f; v[9] = (w22 - w20) / 2.f;
  v[10] = (w13 - w11) / 2.f; v[11] = (w23 - w21) / 2.f;
  v[12] = (w22 + w00 - w20 - w02) / 4.f; v[13] = (w32 + w10 - w30 - w12) / 4.f;
  v[14] = (w23 + w01 - w21 - w03) / 4.f; v[15] = (w33 + w11 - w31 - w13) / 4.f;
  for (int i = 0; i < 16; ++i) {
    float t = 0.f;
    for (int j = 0; j < 16; ++j) t += BICUBIC_M[i * 16 + j] * v[j];
    a[i] = t;
  }
}

/* returns 0 when in bounds, else ERR_INTERP_OOB (w = wx = wy = 0) */
static int interp_bicubic(const uint8_t *img, int rows, int cols, long step, int mult, float xdef,
                          float ydef, float *w, float *wx, float *wy) {
  /* interpolation_class.cpp:82-83 */
  if (xdef > 1.f && ydef > 1.f && xdef < cols - 2.f && ydef < rows - 2.f) {
    int ix = (int)xdef, iy = (int)ydef;
    float a[16];
    bicubic_coeffs(img, step, ix, iy, mult, a);
    float dx = xdef - ix + 1.f;
    float dy = ydef - iy + 1.f;
    float px[4] = {1.f, dx, dx * dx, dx * dx * dx};
    float py[4] = {1.f, dy, dy * dy, dy * dy * dy};
    float rw = 0.f, rx = 0.f, ry = 0.f;
    for (int jk = 0; jk < 4; jk++)
      for (int ik = 0; ik < 4; ik++) {
        int id = jk * 4 + ik;
        rw += a[id] * py[jk] * px[ik];
        if (ik > 0) rx += ik * a[id] * py[jk] * px[ik - 1];
        if (jk > 0) ry += jk * a[id] * py[jk - 1] * px[ik];
      }
    *w = rw; *wx = rx; *wy = ry;
    return 0;
  }
  *w = *wx = *wy = 0.f;
  return ERR_INTERP_OOB;
}

/* interpolation_class.cpp:140-195 + :338-374 */
static int interp_bilinear(const uint8_t *img, int rows, int cols, long step, int mult, float xdef,
                           float ydef, float *w, float *wx, float *wy) {
  if (xdef > 0 && ydef > 0 && xdef < cols - 1 && ydef < rows - 1) {
    int ix = (int)xdef, iy = (int)ydef;
    float w00 = img[step * iy + ix * mult], w01 = img[step * (iy + 1) + ix * mult];
    float w10 = img[step * iy + (ix + 1) * mult], w11 = img[step * (iy + 1) + (ix + 1) * mult];
    float a[4] = {w00, w10 - w00, w01 - w00, w11 - w10 - w01 + w00};
    float dx = xdef - ix, dy = ydef - iy;
    float px[2] = {1.f, dx}, py[2] = {1.f, dy};
    float rw = 0.f, rx = 0.f, ry = 0.f;
    for (int jk = 0; jk < 2; ++jk)
      for (int ik = 0; ik < 2; ++ik) {
        int id = jk * 2 + ik;
        rw += a[id] * py[jk] * px[ik];
        if (ik > 0) rx += a[id] * py[jk];
        if (jk > 0) ry += a[id] * px[ik];
      }
    *w = rw; *wx = rx; *wy = ry;
    return 0;
  }
  *w = *wx = *wy = 0.f;
  return ERR_INTERP_OOB;
}

/* interpolation_class.cpp:197-226 + :376-406 */
static int interp_nearest(const uint8_t *img, int rows, int cols, long step, int mult, int add, float xdef,
                          float ydef, float *w, float *wx, float *wy) {
  if (xdef > 0 && ydef > 0 && xdef < cols - 1 && ydef < rows - 1) {
    int ix = (int)(xdef + 0.5f), iy = (int)(ydef + 0.5f);
    float w00 = img[step * iy + ix * mult + add], w01 = img[step * (iy + 1) + ix * mult + add];
    float w10 = img[step * iy + (ix + 1) * mult + add];
    *w = w00; *wx = w10 - w00; *wy = w01 - w00;
    return 0;
  }
  *w = *wx = *wy = 0.f;
  return ERR_INTERP_OOB;
}

/* ------------------------------------------------------------- one evaluation */

typedef struct {
  const orc_engine *e;
  int level;
  const float *p;
  long first, count;
  float A[ORC_MAXP * ORC_MAXP], b[ORC_MAXP], chi;
  double Ad[ORC_MAXP * ORC_MAXP], bd[ORC_MAXP], chid;
  int error;
} chunk_t;

/* model_class.cpp:48-202: def position and dT/dp rows for one point */
static inline void model_point(int model, const float *p, float x, float y, float cx, float cy,
                               float *xd, float *yd, float *dTx, float *dTy) {
  switch (model) {
  case FM_U:
    *xd = x + p[0]; *yd = y;
    dTx[0] = 1; dTy[0] = 0;
    break;
  case FM_UV:
    *xd = x + p[0]; *yd = y + p[1];
    dTx[0] = 1; dTx[1] = 0; dTy[0] = 0; dTy[1] = 1;
    break;
  case FM_UVQ: {
    float dx = x - cx, dy = y - cy, vx = p[2];
    *xd = x + p[0] - vx * dy;
    *yd = y + p[1] + vx * dx;
    dTx[0] = 1; dTx[1] = 0; dTx[2] = -dy;
    dTy[0] = 0; dTy[1] = 1; dTy[2] = dx;
    break;
  }
  case FM_AFFINE: {
    float dx = x - cx, dy = y - cy;
    *xd = x + p[0] + p[2] * dx + p[3] * dy;
    *yd = y + p[1] + p[4] * dx + p[5] * dy;
    dTx[0] = 1; dTx[1] = 0; dTx[2] = dx; dTx[3] = dy; dTx[4] = 0; dTx[5] = 0;
    dTy[0] = 0; dTy[1] = 1; dTy[2] = 0; dTy[3] = 0; dTy[4] = dx; dTy[5] = dy;
    break;
  }
  case FM_QUAD: { /* our extension, same left-to-right fp32 pattern */
    float dx = x - cx, dy = y - cy;
    *xd = x + p[0] + p[2] * dx + p[3] * dy + 0.5f * p[6] * dx * dx + p[7] * dx * dy +
          0.5f * p[8] * dy * dy;
    *yd = y + p[1] + p[4] * dx + p[5] * dy + 0.5f * p[9] * dx * dx + p[10] * dx * dy +
          0.5f * p[11] * dy * dy;
    float qxx = 0.5f * dx * dx, qxy = dx * dy, qyy = 0.5f * dy * dy;
    for (int i = 0; i < 12; ++i) { dTx[i] = 0; dTy[i] = 0; }
    dTx[0] = 1; dTx[2] = dx; dTx[3] = dy; dTx[6] = qxx; dTx[7] = qxy; dTx[8] = qyy;
    dTy[1] = 1; dTy[4] = dx; dTy[5] = dy; dTy[9] = qxx; dTy[10] = qxy; dTy[11] = qyy;
    break;
  }
  }
}

/* interpolation_class.cpp:671-764 over one thread's contiguous chunk */
static void *chunk_run(void *arg) {
  chunk_t *c = (chunk_t *)arg;
  const orc_engine *e = c->e;
  const int np = e->np, L = c->level;
  const uint8_t *und = e->und.lev[L], *def = e->def.lev[L];
  const int nc = e->colors;
  const long ustep = (long)(e->und.cols / (1 << L)) * nc;
  const int drows = e->def.rows / (1 << L), dcols = e->def.cols / (1 << L);
  const long dstep = (long)(e->def.cols / (1 << L)) * nc;
  const float *xy = e->xy[L] + 2 * c->first;
  const float cx = e->cx[L], cy = e->cy[L];
  float dTx[ORC_MAXP], dTy[ORC_MAXP], H[ORC_MAXP];
  memset(c->A, 0, sizeof(c->A)); memset(c->b, 0, sizeof(c->b)); c->chi = 0.f;
  memset(c->Ad, 0, sizeof(c->Ad)); memset(c->bd, 0, sizeof(c->bd)); c->chid = 0.0;
  c->error = 0;
  for (long i = 0; i < c->count; ++i) {
    float x = xy[2 * i], y = xy[2 * i + 1];
    float xd, yd, w, wx, wy;
    model_point(e->model, c->p, x, y, cx, cy, &xd, &yd, dTx, dTy);
    int und_ix = (int)(x + 0.5f), und_iy = (int)(y + 0.5f);
    for (int col = 0; col < nc; ++col) { /* interpolation_class.cpp:712-750: per-colour loop */
      int err;
      const int mult = nc == 1 ? 1 : nc + col;
      if (e->interp == IM_BICUBIC) err = interp_bicubic(def, drows, dcols, dstep, mult, xd, yd, &w, &wx, &wy);
      else if (e->interp == IM_BILINEAR) err = interp_bilinear(def, drows, dcols, dstep, mult, xd, yd, &w, &wx, &wy);
      else err = interp_nearest(def, drows, dcols, dstep, nc, col, xd, yd, &w, &wx, &wy);
      if (err) c->error = err;
      float und_w = (float)und[ustep * und_iy + (long)und_ix * nc + col];
      float V = und_w - w;
      for (int p = 0; p < np; ++p) H[p] = wx * dTx[p] + wy * dTy[p];
      if (!e->accum_double) {
        c->chi += V * V;
        for (int p1 = 0; p1 < np; ++p1) {
          c->b[p1] += H[p1] * V;
          for (int p2 = p1; p2 < np; ++p2) c->A[p1 * np + p2] += H[p1] * H[p2];
        }
      } else {
        c->chid += (double)(V * V);
        for (int p1 = 0; p1 < np; ++p1) {
          c->bd[p1] += (double)(H[p1] * V);
          for (int p2 = p1; p2 < np; ++p2) c->Ad[p1 * np + p2] += (double)(H[p1] * H[p2]);
        }
      }
    }
  }
  return NULL;
}

/* correlation_class.cpp:131-300 (+ flush_A_B :710-717): split into n_threads contiguous
 * chunks (first N%T get one more), run, fan-in in thread order. Returns the error code. */
static int evaluate(orc_engine *e, int level, const float *p) {
  const int T = e->n_threads, np = e->np;
  const long N = e->npts[level];
  chunk_t *ch = (chunk_t *)malloc(sizeof(chunk_t) * T);
  long per = N / T, first = 0;
  for (int t = 0; t < T; ++t) {
    ch[t].e = e; ch[t].level = level; ch[t].p = p;
    ch[t].count = per + (t < N % T ? 1 : 0);
    ch[t].first = first;
    first += ch[t].count;
  }
  if (e->real_threads && T > 1) {
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * T);
    for (int t = 0; t < T; ++t) pthread_create(&th[t], NULL, chunk_run, &ch[t]);
    for (int t = 0; t < T; ++t) pthread_join(th[t], NULL);
    free(th);
  } else {
    for (int t = 0; t < T; ++t) chunk_run(&ch[t]);
  }
  memset(e->A, 0, sizeof(e->A)); memset(e->b, 0, sizeof(e->b)); e->chi = 0.f;
  int err = 0;
  if (!e->accum_double) {
    for (int t = 0; t < T; ++t) {
      e->chi += ch[t].chi;
      for (int p1 = 0; p1 < np; ++p1) {
        e->b[p1] += ch[t].b[p1];
        for (int p2 = p1; p2 < np; ++p2) e->A[p1 * np + p2] += ch[t].A[p1 * np + p2];
      }
      if (ch[t].error) err = ch[t].error;
    }
  } else {
    double Ad[ORC_MAXP * ORC_MAXP] = {0}, bd[ORC_MAXP] = {0}, chid = 0.0;
    for (int t = 0; t < T; ++t) {
      chid += ch[t].chid;
      for (int p1 = 0; p1 < np; ++p1) {
        bd[p1] += ch[t].bd[p1];
        for (int p2 = p1; p2 < np; ++p2) Ad[p1 * np + p2] += ch[t].Ad[p1 * np + p2];
      }
      if (ch[t].error) err = ch[t].error;
    }
    e->chi = (float)chid;
    for (int i = 0; i < np; ++i) e->b[i] = (float)bd[i];
    for (int i = 0; i < np * np; ++i) e->A[i] = (float)Ad[i];
  }
  free(ch);
  if (err) { e->error_status = 1; e->error_code = err; }
  return err;
}

/* fp64 Gaussian elimination with partial pivoting of the (float-valued) damped system */
static void solve_fp64(const float *A, const float *b, float *x, int n) {
  double M[ORC_MAXP][ORC_MAXP + 1];
  for (int i = 0; i < n; ++i) {
    for (int j = 0; j < n; ++j) M[i][j] = A[i * n + j];
    M[i][n] = b[i];
  }
  for (int k = 0; k < n; ++k) {
    int piv = k;
    for (int i = k + 1; i < n; ++i)
      if (fabs(M[i][k]) > fabs(M[piv][k])) piv = i;
    if (piv != k)
      for (int j = 0; j <= n; ++j) { double t = M[k][j]; M[k][j] = M[piv][j]; M[piv][j] = t; }
    for (int i = k + 1; i < n; ++i) {
      double f = M[i][k] / M[k][k];
      for (int j = k; j <= n; ++j) M[i][j] -= f * M[k][j];
    }
  }
  for (int i = n - 1; i >= 0; --i) {
    double sacc = M[i][n];
    for (int j = i + 1; j < n; ++j) sacc -= M[i][j] * (double)x[j];
    x[i] = (float)(sacc / M[i][i]);
  }
}

void orc_set_solve_double(orc_engine *e, int on) { e->solve_double = on; }

/* correlation_class.cpp:642-688 + solve :719-768. p += dp in place. */
static void compute_model_parameters(orc_engine *e, float *p, float lambda, float scaling) {
  const int np = e->np;
  float *A = e->A, *b = e->b;
  for (int p1 = 0; p1 < np; ++p1) {
    b[p1] *= scaling;
    for (int p2 = p1; p2 < np; ++p2) A[p1 * np + p2] *= scaling;
  }
  for (int p1 = 0; p1 < np; ++p1) {
    for (int p2 = 0; p2 < p1; ++p2) A[p1 * np + p2] = A[p2 * np + p1];
    A[p1 * np + p1] *= (1.f + lambda);
  }
  float dp[ORC_MAXP];
  if (e->solve_double) {
    solve_fp64(A, b, dp, np);
  } else if (e->model == FM_QUAD) {
    /* extension only: Jacobi equilibration S A S y = S b, dp = S y */
    float s[ORC_MAXP], As[ORC_MAXP * ORC_MAXP], bs[ORC_MAXP], y[ORC_MAXP];
    for (int i = 0; i < np; ++i) {
      float d = A[i * np + i];
      s[i] = d > 0.f ? 1.f / sqrtf(d) : 1.f;
    }
    for (int i = 0; i < np; ++i) {
      bs[i] = b[i] * s[i];
      for (int j = 0; j < np; ++j) As[i * np + j] = A[i * np + j] * s[i] * s[j];
    }
    oracle_qr_colpiv_solve(As, bs, y, np);
    for (int i = 0; i < np; ++i) dp[i] = y[i] * s[i];
  } else {
    oracle_qr_colpiv_solve(A, b, dp, np);
  }
  for (int i = 0; i < np; ++i) p[i] += dp[i];
}

int orc_evaluate(orc_engine *e, int level, const float *params, float *A, float *b, float *chi) {
  e->error_status = 0; e->error_code = 0;
  int err = evaluate(e, level, params);
  memcpy(A, e->A, sizeof(float) * e->np * e->np);
  memcpy(b, e->b, sizeof(float) * e->np);
  *chi = e->chi;
  return err;
}

void orc_solve_step(orc_engine *e, const float *A_upper, const float *b, float lambda,
                    float scaling, float *dp) {
  float p[ORC_MAXP] = {0};
  memcpy(e->A, A_upper, sizeof(float) * e->np * e->np);
  memcpy(e->b, b, sizeof(float) * e->np);
  compute_model_parameters(e, p, lambda, scaling);
  memcpy(dp, p, sizeof(float) * e->np);
}

/* --------------------------------------------------------- LM state machine */

static double now_s(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

/* correlation_class.cpp:349-640 */
static void newton_raphson(orc_engine *e, float *mp, orc_result *out) {
  const int np = e->np;
  float last_good[ORC_MAXP], tentative[ORC_MAXP], saved[ORC_MAXP];
  int level_old = 0;
  memset(out->evaluations, 0, sizeof(out->evaluations));
  memset(out->iterations_per_level, 0, sizeof(out->iterations_per_level));
  memset(out->points_per_level, 0, sizeof(out->points_per_level));
  for (int level = e->stop; level >= e->start; level -= e->step) {
    translate_params(e, mp, level_old, level);
    e->error_status = 0; e->error_code = ERR_NONE;
    float lambda = 0.0001f;
    const float min_lambda = 1e-9f, max_lambda = 1e9f;
    e->last_good_chi = FLT_MAX;
    const long N = e->npts[level];
    out->points_per_level[level] = N;
    float scaling = 1.f / ((float)N);
    for (int p = 0; p < np; ++p) last_good[p] = mp[p];

    evaluate(e, level, mp); out->evaluations[level]++;
    if (e->error_status) { /* :413-419 */
      translate_params(e, mp, level, 0);
      return;
    }
    e->chi *= scaling;
    e->last_good_chi = e->chi;
    compute_model_parameters(e, mp, lambda, scaling);
    for (int p = 0; p < np; ++p) saved[p] = mp[p];
    int use_saved = 1;

    for (int iteration = 1; iteration <= e->max_iters + 1; ++iteration) {
      if (iteration > e->max_iters || lambda >= max_lambda) {
        e->error_status = 1; e->error_code = ERR_MAX_ITERS;
        break;
      } else {
        e->reached_iterations = iteration;
        out->iterations_per_level[level] = iteration;
      }
      if (use_saved) {
        for (int p = 0; p < np; ++p) tentative[p] = saved[p];
      } else {
        for (int p = 0; p < np; ++p) mp[p] = last_good[p];
        evaluate(e, level, mp); out->evaluations[level]++;
        e->chi *= scaling;
        if (e->error_status) break;
        compute_model_parameters(e, mp, lambda, scaling);
        for (int p = 0; p < np; ++p) tentative[p] = mp[p];
      }
      for (int p = 0; p < np; ++p) mp[p] = tentative[p];
      evaluate(e, level, mp); out->evaluations[level]++;
      e->chi *= scaling;
      if (e->error_status) break;
      compute_model_parameters(e, mp, fmaxf(lambda * 0.4f, min_lambda), scaling);
      for (int p = 0; p < np; ++p) saved[p] = mp[p];
      float chi = e->chi;
      float delta_chi =
          fabsf((e->last_good_chi - chi) / (fmaxf(e->last_good_chi, chi) + e->precision));
      if (chi <= e->last_good_chi) {
        e->last_good_chi = chi;
        lambda = fmaxf(lambda * 0.4f, min_lambda);
        for (int p = 0; p < np; ++p) last_good[p] = tentative[p];
        use_saved = 1;
      } else {
        lambda = fminf(lambda * 10.0f, max_lambda);
        use_saved = 0;
      }
      if (delta_chi < e->precision) break;
    }
    level_old = level;
  }
  translate_params(e, mp, level_old, 0);
}

static void fill_result(orc_engine *e, const float *p, orc_result *out) {
  memset(out->params, 0, sizeof(out->params));
  for (int i = 0; i < e->np; ++i) out->params[i] = p[i];
  out->chi = e->last_good_chi;
  out->number_of_points = (int)e->npts[0];
  out->iterations = e->reached_iterations;
  out->error_status = e->error_status;
  out->error_code = e->error_code;
  out->und_center_x = e->cx[0];
  out->und_center_y = e->cy[0];
  out->pixel_evaluations = 0;
  for (int l = 0; l < ORC_MAXLEV; ++l)
    out->pixel_evaluations += (double)out->points_per_level[l] * out->evaluations[l];
}

int orc_correlate(orc_engine *e, float *guess_inout, const float *xy, long n, int use_center,
                  float cx, float cy, orc_result *out) {
  double t0 = now_s();
  orc_set_points(e, xy, n, use_center, cx, cy);
  newton_raphson(e, guess_inout, out);
  out->seconds = now_s() - t0;
  fill_result(e, guess_inout, out);
  return out->error_code;
}

int orc_correlate_again(orc_engine *e, float *guess_inout, orc_result *out) {
  double t0 = now_s();
  newton_raphson(e, guess_inout, out);
  out->seconds = now_s() - t0;
  fill_result(e, guess_inout, out);
  return out->error_code;
}

/* ------------------------------------------------------- pixel-list builders */

/* manager_class.cpp:1596-1614: x outer, y inner, both ends inclusive */
long orc_rect_points(int x0, int y0, int x1, int y1, float *out, long cap) {
  long m = 0;
  for (int ix = x0; ix <= x1; ++ix)
    for (int iy = y0; iy <= y1; ++iy) {
      if (m < cap) { out[2 * m] = (float)ix; out[2 * m + 1] = (float)iy; }
      ++m;
    }
  return m;
}

/* manager_class.cpp:816-940 (serial order: i outer, j inner). `as` = angular subdivisions. */
long orc_annulus_points(float r, float dr, float a, float da, float cx, float cy, int as,
                        float *out, long cap) {
  int x0, y0, x1, y1;
  float c00x = 0, c01x = 0, c10x = 0, c11x = 0, c00y = 0, c01y = 0, c10y = 0, c11y = 0;
  if (as <= 0) return -1;
  if (as == 1) {
    x0 = cx - (r + dr); x1 = cx + (r + dr);
    y0 = cy - (r + dr); y1 = cy + (r + dr);
  } else {
    float sin0 = (float)sin(a), cos0 = (float)cos(a);
    float sin1 = (float)sin(a + da), cos1 = (float)cos(a + da);
    float sin2 = (float)sin(a + da / 2.f), cos2 = (float)cos(a + da / 2.f);
    c00x = cx + (r)*cos0; c01x = cx + (r)*cos1;
    c10x = cx + (r + dr) * cos0 * 1.2f; c11x = cx + (r + dr) * cos1 * 1.2f;
    c00y = cy + (r)*sin0; c01y = cy + (r)*sin1;
    c10y = cy + (r + dr) * sin0 * 1.2f; c11y = cy + (r + dr) * sin1 * 1.2f;
    float arc_x = cx + (r + dr) * cos2, arc_y = cy + (r + dr) * sin2;
    x0 = fminf(arc_x, fminf(fminf(c00x, c01x), fminf(c10x, c11x)));
    x1 = fmaxf(arc_x, fmaxf(fmaxf(c00x, c01x), fmaxf(c10x, c11x)));
    y0 = fminf(arc_y, fminf(fminf(c00y, c01y), fminf(c10y, c11y)));
    y1 = fmaxf(arc_y, fmaxf(fmaxf(c00y, c01y), fmaxf(c10y, c11y)));
  }
  float ro2 = (r + dr) * (r + dr);
  float ri2 = r * r;
  long m = 0;
  for (float i = x0; i < x1; ++i)
    for (int j = y0; j < y1; ++j) {
      float r2 = (i - cx) * (i - cx) + (j - cy) * (j - cy);
      if (r2 > ri2 && r2 < ro2) {
        float cross1 = (c11x - i) * (c01y - c11y) - (c11y - j) * (c01x - c11x);
        float cross2 = (c00x - i) * (c10y - c00y) - (c00y - j) * (c10x - c00x);
        if (cross1 * cross2 > 0 || as == 1) {
          if (m < cap) { out[2 * m] = i; out[2 * m + 1] = (float)j; }
          ++m;
        }
      }
    }
  return m;
}

/* ---- blob: polygon_class.cpp (O'Rourke ear clipping + half-open scanline fill) ---- */

typedef struct {
  float x, y;
  int ear;
  int next, prev;
} vtx_t;

typedef struct {
  vtx_t *v;
  int head, count;
} poly_t;

static float area2(const vtx_t *v, int a, int b, int c) { /* :49-57 */
  return (v[b].x - v[a].x) * (v[c].y - v[a].y) - (v[c].x - v[a].x) * (v[b].y - v[a].y);
}
static int p_left(const vtx_t *v, int a, int b, int c) { return area2(v, a, b, c) > 0.f; }
static int p_lefton(const vtx_t *v, int a, int b, int c) { return area2(v, a, b, c) >= 0.f; }
static int p_collinear(const vtx_t *v, int a, int b, int c) { return area2(v, a, b, c) == 0.f; }

static int p_intersect_prop(const vtx_t *v, int a, int b, int c, int d) { /* :109-118 */
  if (p_collinear(v, a, b, c) || p_collinear(v, a, b, d) || p_collinear(v, b, d, a) ||
      p_collinear(v, c, d, b))
    return 0;
  return (!p_left(v, a, b, c) ^ !p_left(v, a, b, d)) && (!p_left(v, c, d, a) ^ !p_left(v, c, d, b));
}
static int p_between(const vtx_t *v, int a, int b, int c) { /* :120-139 */
  if (!p_collinear(v, a, b, c)) return 0;
  if (v[a].x != v[b].x)
    return ((v[a].x <= v[c].x) && (v[c].x <= v[b].x)) || ((v[a].x >= v[c].x) && (v[c].x >= v[b].x));
  return ((v[a].y <= v[c].y) && (v[c].y <= v[b].y)) || ((v[a].y >= v[c].y) && (v[c].y >= v[b].y));
}
static int p_intersect(const vtx_t *v, int a, int b, int c, int d) { /* :141-152 */
  if (p_intersect_prop(v, a, b, c, d)) return 1;
  if (p_between(v, a, b, c) || p_between(v, a, b, d) || p_between(v, c, d, a) ||
      p_between(v, c, d, b))
    return 1;
  return 0;
}
static int p_diagonal_ie(const poly_t *P, int a, int b) { /* :154-173 */
  const vtx_t *v = P->v;
  int c = P->head;
  do {
    int c1 = v[c].next;
    if ((c != a) && (c1 != a) && (c != b) && (c1 != b) && p_intersect(v, a, b, c, c1)) return 0;
    c = v[c].next;
  } while (c != P->head);
  return 1;
}
static int p_in_cone(const poly_t *P, int a, int b) { /* :175-187 */
  const vtx_t *v = P->v;
  int a1 = v[a].next, a0 = v[a].prev;
  if (p_lefton(v, a, a1, a0)) return p_left(v, a, b, a0) && p_left(v, b, a, a1);
  return !(p_lefton(v, a, b, a1) && p_lefton(v, b, a, a0));
}
static int p_diagonal(const poly_t *P, int a, int b) { /* :189-191 */
  return p_in_cone(P, a, b) && p_in_cone(P, b, a) && p_diagonal_ie(P, a, b);
}
static float p_area_poly2(const poly_t *P) { /* :68-81 */
  const vtx_t *v = P->v;
  float sum = 0.f;
  int a = v[P->head].next;
  do {
    sum += area2(v, P->head, a, v[a].next);
    a = v[a].next;
  } while (v[a].next != P->head);
  return sum;
}
static int p_simple_loop(const poly_t *P) { /* :195-222 */
  const vtx_t *v = P->v;
  if (P->count < 4) return 1;
  int ol = P->head, orr;
  do {
    orr = v[ol].next;
    int il = v[orr].next, ir;
    do {
      ir = v[il].next;
      if (p_intersect(v, ol, orr, il, ir)) return 0;
      il = ir;
    } while (il != P->head && il != v[ol].prev);
    ol = orr;
  } while (ol != v[v[P->head].prev].prev);
  return 1;
}

typedef struct { float *out; long cap, m; } sink_t;
static void sink_push(sink_t *s, int i, int j) {
  if (s->m < s->cap) { s->out[2 * s->m] = (float)i; s->out[2 * s->m + 1] = (float)j; }
  s->m++;
}
static int p_line(float x1, float y1, float x2, float y2, float *dxdy, float *x0) { /* :405-416 */
  float den = y2 - y1;
  if (den != 0) {
    *dxdy = (x2 - x1) / den;
    *x0 = x1 - *dxdy * y1;
    return 0;
  }
  return 1;
}
/* :339-403; v1 and v2 share y */
static void flat_triangle(sink_t *s, float x1, float y1, float x2, float y2, float x3, float y3) {
  (void)y2;
  int dy = (int)(floor(y3) - floor(y1));
  int dx = (int)(floor(x2) - floor(x1));
  if (dx == 0 || dy == 0) return;
  float xs, ys, xb, yb;
  if (dx > 0) { xs = x1; ys = y1; xb = x2; yb = y2; }
  else { xs = x2; ys = y2; xb = x1; yb = y1; }
  float ds = 0, dbg = 0, x0s = 0, x0b = 0;
  p_line(xs, ys, x3, y3, &ds, &x0s);
  p_line(xb, yb, x3, y3, &dbg, &x0b);
  int j0, j1;
  if (dy > 0) { j0 = (int)ceil(y1); j1 = (int)ceil(y3); }
  else { j0 = (int)ceil(y3); j1 = (int)ceil(y1); }
  for (int j = j0; j < j1; ++j) {
    int i0 = (int)ceilf(ds * (float)j + x0s);
    int i1 = (int)ceilf(dbg * (float)j + x0b);
    for (int i = i0; i < i1; ++i) sink_push(s, i, j);
  }
}
/* :283-337. (The three flatTrianglePoints calls at :285-298 discard their result.) */
static void triangle_points(sink_t *s, const vtx_t *a, const vtx_t *b, const vtx_t *c) {
  const vtx_t *ymax, *ymid, *ymin;
  if (a->y > b->y) {
    if (b->y > c->y) { ymax = a; ymid = b; ymin = c; }
    else if (c->y > a->y) { ymax = c; ymid = a; ymin = b; }
    else { ymax = a; ymid = c; ymin = b; }
  } else {
    if (a->y > c->y) { ymax = b; ymid = a; ymin = c; }
    else if (c->y > b->y) { ymax = c; ymid = b; ymin = a; }
    else { ymax = b; ymid = c; ymin = a; }
  }
  float dxdy, x0;
  if (p_line(ymin->x, ymin->y, ymax->x, ymax->y, &dxdy, &x0)) return;
  float newY = ymid->y;
  float newX = dxdy * newY + x0;
  flat_triangle(s, ymid->x, ymid->y, newX, newY, ymax->x, ymax->y);
  flat_triangle(s, ymid->x, ymid->y, newX, newY, ymin->x, ymin->y);
}

/* Triangulates (polygon_class.cpp:224-281) and rasterises (:418-429). Returns the number of
 * inside points, -1 for a self-intersecting contour (error_bad_domain), -2 if ear clipping
 * stalls (the reference would loop forever). tri_out (optional, 6 floats per triangle,
 * capacity tri_cap triangles) receives the triangles in emission order; *n_tri their count. */
long orc_blob_points(const float *contour, int n, float *out, long cap, float *tri_out,
                     int tri_cap, int *n_tri) {
  poly_t P;
  P.v = (vtx_t *)malloc(sizeof(vtx_t) * (size_t)(n > 0 ? n : 1));
  P.count = n; P.head = 0;
  for (int i = 0; i < n; ++i) {
    P.v[i].x = contour[2 * i]; P.v[i].y = contour[2 * i + 1];
    P.v[i].ear = 0;
    P.v[i].next = (i + 1) % n; P.v[i].prev = (i + n - 1) % n;
  }
  if (n_tri) *n_tri = 0;
  if (n < 3 || !p_simple_loop(&P)) { free(P.v); return -1; }
  if (p_area_poly2(&P) < 0) { /* reOrientPoly :83-97 */
    for (int i = 0; i < n; ++i) { int t = P.v[i].prev; P.v[i].prev = P.v[i].next; P.v[i].next = t; }
  }
  vtx_t *v = P.v;
  { /* earInit :37-47 */
    int v1 = P.head;
    do {
      v[v1].ear = p_diagonal(&P, v[v1].prev, v[v1].next);
      v1 = v[v1].next;
    } while (v1 != P.head);
  }
  sink_t s = {out, cap, 0};
  int nt = 0;
#define EMIT(A_, B_, C_)                                                        \
  do {                                                                          \
    if (tri_out && nt < tri_cap) {                                              \
      float *t = tri_out + 6 * nt;                                              \
      t[0] = v[A_].x; t[1] = v[A_].y; t[2] = v[B_].x; t[3] = v[B_].y;           \
      t[4] = v[C_].x; t[5] = v[C_].y;                                           \
    }                                                                           \
    ++nt;                                                                       \
    triangle_points(&s, &v[A_], &v[B_], &v[C_]);                                \
  } while (0)
  while (P.count > 3) {
    int v2 = P.head, found = 0;
    do {
      if (v[v2].ear) {
        int v3 = v[v2].next, v4 = v[v3].next, v1 = v[v2].prev, v0 = v[v1].prev;
        EMIT(v1, v2, v3);
        v[v1].ear = p_diagonal(&P, v0, v3);
        v[v3].ear = p_diagonal(&P, v1, v4);
        v[v1].next = v3; v[v3].prev = v1;
        P.head = v3; P.count--;
        found = 1;
        break;
      }
      v2 = v[v2].next;
    } while (v2 != P.head);
    if (!found) { free(P.v); return -2; }
  }
  {
    int v2 = P.head, v1 = v[v2].prev, v3 = v[v2].next;
    EMIT(v1, v2, v3);
  }
#undef EMIT
  if (n_tri) *n_tri = nt;
  free(P.v);
  return s.m;
}

/* pyramid_class.cpp:325-347: the centre the CPU engine derives from a list */
void orc_seq_mean_center(const float *xy, long n, float *cx, float *cy) {
  float sx = 0.f, sy = 0.f;
  for (long i = 0; i < n; ++i) { sx += xy[2 * i]; sy += xy[2 * i + 1]; }
  *cx = sx / (float)n; *cy = sy / (float)n;
}

/* parameters.cpp:55-58 */
float orc_best_rotation(const float *p) {
  return (float)atan2((double)(p[4] - p[3]), (double)(p[2] + p[5] + 2.f));
}

int orc_result_size(void) { return (int)sizeof(orc_result); }

// oracle/ref_driver.cpp -- TEST INFRASTRUCTURE (oracle). Never linked into the product.
//
// Thin extern "C" driver around the UNMODIFIED reference CPU engine, compiled from
// the sources where they lie under /root/reference (see oracle/Makefile):
//   correlation_class.cpp interpolation_class.cpp model_class.cpp pyramid_class.cpp
//   parameters.cpp polygon_class.cpp
// against the three shim headers in oracle/shim (cv::Mat storage, Q_DECLARE_METATYPE,
// Eigen colPivHouseholderQr stand-in). The output lives only in oracle/_ref/.
//
// It replays what managerClass does around CorrelationClass
// (manager_class.cpp:1339-1342 construction, :1406-1407 image setters,
//  :438-441 / :700-702 / :1141-1143 Newton_Raphson overloads, :2319-2331 getters)
// so tests can (1) pin the C restatement (oracle/dic_oracle.c) bit-for-bit and
// (2) generate the golden fixtures under tests/golden/.
//
// The only liberty taken: this TU (and only this TU) sees the classes' private
// members, to read A, b, chi after a single evaluation and the pyramid levels.
// Access specifiers do not change object layout with GCC, and the reference TUs
// themselves are compiled untouched.
// Standard and shim headers first, so the keyword trick below only ever touches
// the reference's own class declarations.
#include <algorithm>
#include <assert.h>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <ctime>
#include <future>
#include <iostream>
#include <limits>
#include <math.h>
#include <memory>
#include <stdio.h>
#include <thread>
#include <utility>
#include <vector>
#include "opencv2/core/core.hpp"
#include <Dense>
#include <QMetaType>

#define private public
#define protected public
#include "correlation_class.hpp"
#include "polygon_class.h"
#undef private
#undef protected

#include <chrono>
#include <cstdint>
#include <cstring>
#include <vector>

extern "C" {

struct ref_result {
  float params[12];
  float chi;
  int number_of_points;
  int iterations;
  int error_code;
  int error_status;
  float und_center_x;
  float und_center_y;
  double seconds; // wall time of the Newton_Raphson call alone
};

struct ref_engine {
  CorrelationClass *corr{nullptr};
  cv::Mat und, def, nxt;
  std::vector<float> xy; // level-0 list: owned here, borrowed by the pyramid
  int n_params{0};
  int colors{1};
};

ref_engine *ref_create(int n_threads, int interpolation_model, int fitting_model,
                       float precision, int max_iters, int pyr_start,
                       int pyr_step, int pyr_stop) {
  ref_engine *e = new ref_engine;
  // manager_class.cpp:1339-1342: (allocated_points, colors, interp, model, threads,
  // precision, max_iters, start, step, stop); monochrome => 1 colour.
  e->corr = new CorrelationClass(
      1, 1, (interpolationModelEnum)interpolation_model,
      (fittingModelEnum)fitting_model, n_threads, precision, max_iters,
      pyr_start, pyr_step, pyr_stop);
  e->n_params =
      ModelClass::get_number_of_model_parameters((fittingModelEnum)fitting_model);
  return e;
}

// colour images: number_of_colors = 3 (manager_class.cpp:99-109), interleaved 8-bit channels
ref_engine *ref_create_color(int n_threads, int interpolation_model, int fitting_model, float precision,
                             int max_iters, int pyr_start, int pyr_step, int pyr_stop, int colors) {
  ref_engine *e = new ref_engine;
  e->colors = colors == 3 ? 3 : 1;
  e->corr = new CorrelationClass(1, e->colors, (interpolationModelEnum)interpolation_model,
                                 (fittingModelEnum)fitting_model, n_threads, precision, max_iters, pyr_start,
                                 pyr_step, pyr_stop);
  e->n_params = ModelClass::get_number_of_model_parameters((fittingModelEnum)fitting_model);
  return e;
}

void ref_destroy(ref_engine *e) {
  if (!e) return;
  // CorrelationClass / Pyramid_class destructors free levels >= 1 only.
  // Level-0 list belongs to us (reference: to the manager).
  delete e->corr;
  delete e;
}

static cv::Mat make_mat(const uint8_t *img, int rows, int cols, int colors = 1) {
  cv::Mat m(rows, cols, colors == 3 ? CV_8UC3 : CV_8U);
  std::memcpy(m.data, img, (size_t)rows * cols * colors);
  return m;
}

void ref_set_und_image(ref_engine *e, const uint8_t *img, int rows, int cols) {
  e->und = make_mat(img, rows, cols, e->colors);
  e->corr->set_undeformed_image(e->und);
}
void ref_set_def_image(ref_engine *e, const uint8_t *img, int rows, int cols) {
  e->def = make_mat(img, rows, cols, e->colors);
  e->corr->set_deformed_image(e->def);
}
void ref_set_nxt_image(ref_engine *e, const uint8_t *img, int rows, int cols) {
  e->nxt = make_mat(img, rows, cols, e->colors);
  e->corr->set_next_image(e->nxt);
}
void ref_und_from_def(ref_engine *e) {
  e->und = e->def;
  e->corr->set_und_image_from_def();
}
void ref_def_from_nxt(ref_engine *e) {
  e->def = e->nxt;
  e->corr->set_def_image_from_nxt();
}

static void fill_result(ref_engine *e, const float *p, ref_result *out) {
  std::memset(out->params, 0, sizeof(out->params));
  for (int i = 0; i < e->n_params; ++i) out->params[i] = p[i];
  out->chi = e->corr->get_chi();
  out->number_of_points = e->corr->get_number_of_points();
  out->iterations = e->corr->get_iterations();
  out->error_status = e->corr->get_error_status() ? 1 : 0;
  out->error_code = (int)e->corr->get_error_code();
  out->und_center_x = e->corr->get_und_x_center();
  out->und_center_y = e->corr->get_und_y_center();
}

// use_center != 0: the "blob with known center" overload (manager rect path,
// manager_class.cpp:438-441); else centre = sequential fp32 mean of the list
// (annulus/blob path, :700-702, :1141-1143 -> pyramid_class.cpp:325-347).
int ref_correlate(ref_engine *e, float *guess_inout, const float *xy, int n,
                  int use_center, float cx, float cy, ref_result *out) {
  e->xy.assign(xy, xy + 2 * (size_t)n);
  auto t0 = std::chrono::steady_clock::now();
  float *p;
  if (use_center)
    p = e->corr->Newton_Raphson(guess_inout, n, cx, cy, e->xy.data());
  else
    p = e->corr->Newton_Raphson(guess_inout, n, e->xy.data());
  auto t1 = std::chrono::steady_clock::now();
  fill_result(e, p, out);
  out->seconds = std::chrono::duration<double>(t1 - t0).count();
  return out->error_code;
}

// Same point set, new guess only (correlation_class.cpp:349).
int ref_correlate_again(ref_engine *e, float *guess_inout, ref_result *out) {
  auto t0 = std::chrono::steady_clock::now();
  float *p = e->corr->Newton_Raphson(guess_inout);
  auto t1 = std::chrono::steady_clock::now();
  fill_result(e, p, out);
  out->seconds = std::chrono::duration<double>(t1 - t0).count();
  return out->error_code;
}

// ---- single-evaluation access (private members) -------------------------------

// Installs a point set exactly as Newton_Raphson's overloads do
// (correlation_class.cpp:306-343) without running the solver.
void ref_set_points(ref_engine *e, const float *xy, int n, int use_center,
                    float cx, float cy) {
  CorrelationClass *c = e->corr;
  e->xy.assign(xy, xy + 2 * (size_t)n);
  if (c->allocated_points < n) {
    c->allocated_points = n;
    c->delete_point_dependent_arrays();
    c->allocate_point_dependent_arrays();
  }
  c->pyramid.set_xy_positions(e->xy.data(), n);
  if (use_center)
    c->pyramid.set_und_center(cx, cy);
  else
    c->pyramid.set_und_center();
}

int ref_level_num_points(ref_engine *e, int level) {
  return e->corr->pyramid.get_number_of_points(level);
}
void ref_level_points(ref_engine *e, int level, float *xy_out) {
  int n = e->corr->pyramid.get_number_of_points(level);
  std::memcpy(xy_out, e->corr->pyramid.get_xy_positions(level),
              sizeof(float) * 2 * (size_t)n);
}
void ref_level_center(ref_engine *e, int level, float *cx, float *cy) {
  e->corr->pyramid.get_und_center(*cx, *cy, level);
}
// which: 0 = und, 1 = def
void ref_pyramid_level(ref_engine *e, int which, int level, uint8_t *out,
                       int *rows, int *cols) {
  Pyramid_class &p = e->corr->pyramid;
  ImageType t = which == 0 ? imageType_und : imageType_def;
  int r = p.get_rows(level, t), c = p.get_cols(level, t);
  *rows = r;
  *cols = c;
  if (out)
    std::memcpy(out, which == 0 ? p.get_und_ptr(level) : p.get_def_ptr(level),
                (size_t)r * c * e->colors);
}

// One evaluation (flush_A_B + apply_model_and_interpolate,
// correlation_class.cpp:410-411) at `params` given in LEVEL units.
// Returns raw (unscaled) upper-triangular A (row-major n x n), b, chi.
int ref_evaluate(ref_engine *e, int level, const float *params, float *A,
                 float *b, float *chi) {
  CorrelationClass *c = e->corr;
  int n = e->n_params;
  std::vector<float> p(params, params + n);
  c->model_parameters = p.data();
  c->error_status = false;
  c->error_code = error_none;
  c->flush_A_B();
  c->apply_model_and_interpolate(level, true);
  std::memcpy(A, c->mat_A, sizeof(float) * n * n);
  std::memcpy(b, c->vec_B, sizeof(float) * n);
  *chi = c->chi;
  c->model_parameters = nullptr;
  return c->error_status ? (int)c->error_code : 0;
}

// scale + mirror + damp + solve (correlation_class.cpp:642-688) on caller data.
void ref_solve_step(ref_engine *e, const float *A_upper, const float *b,
                    float lambda, float scaling, float *dp) {
  CorrelationClass *c = e->corr;
  int n = e->n_params;
  std::vector<float> p(n, 0.f);
  c->model_parameters = p.data();
  std::memcpy(c->mat_A, A_upper, sizeof(float) * n * n);
  std::memcpy(c->vec_B, b, sizeof(float) * n);
  c->compute_model_parameters(lambda, scaling);
  for (int i = 0; i < n; ++i) dp[i] = p[i];
  c->model_parameters = nullptr;
}

// ---- blob rasterisation (polygon_class.cpp) ------------------------------------
// Returns the number of inside points (writes at most cap of them), or -1 when
// the contour self-intersects (polygon_class.cpp:225-229 -> error_bad_domain).
long ref_blob_points(const float *contour_xy, int n_vertices, float *xy_out,
                     long cap) {
  v_points contour(n_vertices);
  for (int i = 0; i < n_vertices; ++i)
    contour[i] = std::make_pair(contour_xy[2 * i], contour_xy[2 * i + 1]);
  polygonBlob_class polygon(contour);
  if (polygon.getError()) return -1;
  v_points pts = polygon.getInsidePoints();
  long m = (long)pts.size();
  for (long i = 0; i < m && i < cap; ++i) {
    xy_out[2 * i] = pts[i].first;
    xy_out[2 * i + 1] = pts[i].second;
  }
  return m;
}

float ref_best_rotation(float *p) { return best_rotation_UVUxUyVxVy(p); }

} // extern "C"

// dic_host_c.cpp -- C entry point (for ctypes / any FFI) and command-line front end of the
// headless host (dic_manager.hpp). Links against libdic_b200.so only.
//
//   libdic_host.so : dic_host_run(...)                       (tests, bench.py --workload c3)
//   dic_headless   : dic_headless --rect x0 y0 x1 y1 hs vs --out report.csv f0.pgm f1.pgm ...
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>

#include "dic_manager.hpp"

extern "C" {

struct dic_host_config {
  int domain_type;            // 0 rect, 1 annulus, 2 blob
  float rect[4];              // x_begin y_begin x_end y_end
  int subdivisions[2];        // horizontal, vertical | radial, angular
  float annulus[4];           // x_center y_center r_inside r_outside
  const float *contour_xy;    // blob
  int n_contour;
  int model, interpolation;
  int pyramid[3];
  float precision;
  int max_iters;
  int deformation_description; // 0 strict Lagrangian, 1 Lagrangian, 2 Eulerian
  int reference_image;         // 0 first, 1 previous
  float global_initial_guess[DIC_MAX_PARAMS];
  int arith_mode;
  int batch_sectors;
  int device;
  int error_handling_mode;     // 0 stopAll (GUI default), 1 stopFrame, 2 continue (enums.hpp:80-85)
};

struct dic_host_row { // one CSV row, numerically
  int frame, sector;
  float und_center_x, und_center_y, def_center_x, def_center_y, def_angle;
  float params[DIC_MAX_PARAMS], initial_guess[DIC_MAX_PARAMS];
  float chi;
  int number_of_points, iterations, error_code;
  float und_global_center_x, und_global_center_y, def_global_center_x, def_global_center_y, def_global_angle, und_angle;
};

static dic_host::Config to_config(const dic_host_config *c) {
  dic_host::Config k;
  k.domain_type = (dic_host::domainEnum)c->domain_type;
  k.x_begin = c->rect[0]; k.y_begin = c->rect[1]; k.x_end = c->rect[2]; k.y_end = c->rect[3];
  if (c->domain_type == 0) { k.horizontal_subdivisions = c->subdivisions[0]; k.vertical_subdivisions = c->subdivisions[1]; }
  else { k.radial_subdivisions = c->subdivisions[0]; k.angular_subdivisions = c->subdivisions[1]; }
  k.x_center = c->annulus[0]; k.y_center = c->annulus[1]; k.r_inside = c->annulus[2]; k.r_outside = c->annulus[3];
  for (int i = 0; i < c->n_contour; ++i) k.xy_contour.push_back(std::make_pair(c->contour_xy[2 * i], c->contour_xy[2 * i + 1]));
  k.model = (fittingModelEnum)c->model; k.interpolation = (interpolationModelEnum)c->interpolation;
  k.pyramid_start = c->pyramid[0]; k.pyramid_step = c->pyramid[1]; k.pyramid_stop = c->pyramid[2];
  k.precision = c->precision; k.max_iters = c->max_iters;
  k.deformationDescription = (deformationDescriptionEnum)c->deformation_description;
  k.referenceImage = (referenceImageEnum)c->reference_image;
  memcpy(k.global_initial_guess, c->global_initial_guess, sizeof(k.global_initial_guess));
  k.arith_mode = c->arith_mode; k.batch_sectors = c->batch_sectors != 0;
  k.error_handling_mode = (errorHandlingModeEnum)c->error_handling_mode;
  return k;
}

// Runs the whole multi-frame correlation. csv: receives the report (NUL terminated, truncated to
// csv_cap); rows (optional, capacity rows_cap) receives the last frame's sectors... every frame's
// sectors in order. Returns 0 / 1 = the manager's error flag, negative on set-up failure.
int dic_host_run(const dic_host_config *c, const uint8_t *const *frames, int n_frames, int rows, int cols,
                 char *csv, long long csv_cap, long long *csv_needed, double *seconds, dic_host_row *out_rows,
                 int rows_cap, int *rows_written) {
  try {
    dic_host::Config k = to_config(c);
    dic_host::HeadlessManager m(k, c->device);
    std::vector<const uint8_t *> f(frames, frames + n_frames);
    // per-frame rows are captured by re-running the report parser below; timing covers the loop only
    auto t0 = std::chrono::steady_clock::now();
    bool err = m.perform_multiframe_correlation(f, rows, cols);
    auto t1 = std::chrono::steady_clock::now();
    if (seconds) *seconds = std::chrono::duration<double>(t1 - t0).count();
    std::string rep = m.report();
    if (csv_needed) *csv_needed = (long long)rep.size() + 1;
    if (csv && csv_cap > 0) {
      size_t n = std::min<size_t>(rep.size(), (size_t)csv_cap - 1);
      memcpy(csv, rep.data(), n);
      csv[n] = 0;
    }
    int w = 0;
    if (out_rows) { // final state of every sector (exact floats, not the 6-digit CSV)
      const auto &rs = m.results();
      for (size_t s = 0; s < rs.size() && w < rows_cap; ++s, ++w) {
        dic_host_row &o = out_rows[w];
        o.frame = m.frames_done() - 1; o.sector = (int)s;
        o.und_global_center_x = rs[s].und_global_center_x; o.und_global_center_y = rs[s].und_global_center_y;
        o.def_global_center_x = rs[s].def_global_center_x; o.def_global_center_y = rs[s].def_global_center_y;
        o.def_global_angle = rs[s].def_global_angle; o.und_angle = rs[s].und_angle;
        o.und_center_x = rs[s].und_center_x; o.und_center_y = rs[s].und_center_y;
        o.def_center_x = rs[s].def_center_x; o.def_center_y = rs[s].def_center_y; o.def_angle = rs[s].def_angle;
        memcpy(o.params, rs[s].resulting_parameters, sizeof(o.params));
        memcpy(o.initial_guess, rs[s].initial_guess, sizeof(o.initial_guess));
        o.chi = rs[s].chi; o.number_of_points = rs[s].number_of_points; o.iterations = rs[s].iterations;
        o.error_code = (int)rs[s].error_code;
      }
    }
    if (rows_written) *rows_written = w;
    return err ? 1 : 0;
  } catch (const std::exception &ex) {
    fprintf(stderr, "dic_host_run: %s\n", ex.what());
    return -1;
  }
}

// managerClass::deformPoints (manager_class.cpp:2527-2600) for FFI callers: n points (x, y interleaved)
// moved by `params` of `model` about the centre (cx, cy). No GPU involved.
int dic_host_deform_points(int model, const float *params, float cx, float cy, const float *xy_in, int n,
                           float *xy_out) {
  if (!params || !xy_in || !xy_out || n < 0 || model < fm_U || model > fm_UVUxUyVxVyQuad) return -1;
  for (int i = 0; i < n; ++i)
    dic_host::distort_point(model, xy_in[2 * i], xy_in[2 * i + 1], cx, cy, params, xy_out[2 * i], xy_out[2 * i + 1]);
  return 0;
}

} // extern "C"

#ifdef DIC_HEADLESS_MAIN
static bool read_pgm(const char *path, std::vector<uint8_t> &img, int &rows, int &cols) {
  std::ifstream f(path, std::ios::binary);
  std::string magic;
  int maxv = 0;
  f >> magic;
  auto skip = [&] { while (f >> std::ws && f.peek() == '#') { std::string l; std::getline(f, l); } };
  skip(); f >> cols; skip(); f >> rows; skip(); f >> maxv;
  f.get();
  if (!f || magic != "P5" || maxv != 255) return false;
  img.resize((size_t)rows * cols);
  f.read(reinterpret_cast<char *>(img.data()), (std::streamsize)img.size());
  return (bool)f;
}

int main(int argc, char **argv) {
  dic_host_config c;
  memset(&c, 0, sizeof(c));
  c.model = 3; c.interpolation = 2; c.pyramid[0] = 0; c.pyramid[1] = 1; c.pyramid[2] = 2;
  c.precision = 1e-3f; c.max_iters = 50; c.deformation_description = 2; c.reference_image = 0;
  c.subdivisions[0] = c.subdivisions[1] = 1;
  std::string out = "report.csv";
  std::vector<float> contour;
  std::vector<std::string> files;
  for (int i = 1; i < argc; ++i) {
    std::string a = argv[i];
    auto f = [&](int k) { return (float)atof(argv[i + k]); };
    if (a == "--rect" && i + 6 < argc) { c.domain_type = 0; for (int k = 0; k < 4; ++k) c.rect[k] = f(k + 1); c.subdivisions[0] = atoi(argv[i + 5]); c.subdivisions[1] = atoi(argv[i + 6]); i += 6; }
    else if (a == "--annulus" && i + 6 < argc) { c.domain_type = 1; for (int k = 0; k < 4; ++k) c.annulus[k] = f(k + 1); c.subdivisions[0] = atoi(argv[i + 5]); c.subdivisions[1] = atoi(argv[i + 6]); i += 6; }
    else if (a == "--blob") { c.domain_type = 2; while (i + 2 < argc && argv[i + 1][0] != '-' ) { contour.push_back(f(1)); contour.push_back(f(2)); i += 2; } }
    else if (a == "--model") c.model = atoi(argv[++i]);
    else if (a == "--pyramid" && i + 3 < argc) { for (int k = 0; k < 3; ++k) c.pyramid[k] = atoi(argv[i + 1 + k]); i += 3; }
    else if (a == "--precision") c.precision = (float)atof(argv[++i]);
    else if (a == "--max-iters") c.max_iters = atoi(argv[++i]);
    else if (a == "--lagrangian") c.deformation_description = 1;
    else if (a == "--previous") c.reference_image = 1;
    else if (a == "--fast") c.arith_mode = 1;
    else if (a == "--batch") c.batch_sectors = 1;
    else if (a == "--on-error" && i + 1 < argc) { std::string m = argv[++i]; c.error_handling_mode = m == "stopFrame" ? 1 : m == "continue" ? 2 : 0; }
    else if (a == "--strict-lagrangian") c.deformation_description = 0;
    else if (a == "--out") out = argv[++i];
    else files.push_back(a);
  }
  c.contour_xy = contour.data(); c.n_contour = (int)contour.size() / 2;
  if (files.size() < 2) { fprintf(stderr, "usage: dic_headless [--rect x0 y0 x1 y1 hs vs | --annulus cx cy ri ro rs as | --blob x y ...] [--model m] [--pyramid a b c] --out report.csv f0.pgm f1.pgm ...\n"); return 2; }
  std::vector<std::vector<uint8_t>> imgs(files.size());
  std::vector<const uint8_t *> ptrs;
  int rows = 0, cols = 0;
  for (size_t k = 0; k < files.size(); ++k) {
    int r, q;
    if (!read_pgm(files[k].c_str(), imgs[k], r, q) || (k && (r != rows || q != cols))) { fprintf(stderr, "cannot read %s\n", files[k].c_str()); return 2; }
    rows = r; cols = q; ptrs.push_back(imgs[k].data());
  }
  long long need = 0; double secs = 0;
  std::vector<char> csv(1 << 20);
  int rc = dic_host_run(&c, ptrs.data(), (int)ptrs.size(), rows, cols, csv.data(), (long long)csv.size(), &need, &secs, nullptr, 0, nullptr);
  if (need > (long long)csv.size()) { csv.resize(need); rc = dic_host_run(&c, ptrs.data(), (int)ptrs.size(), rows, cols, csv.data(), need, &need, &secs, nullptr, 0, nullptr); }
  std::ofstream(out) << csv.data();
  fprintf(stderr, "dic_headless: %zu frame pairs in %.3f s (%.1f frames/s), error flag %d, report -> %s\n", ptrs.size() - 1, secs, (ptrs.size() - 1) / secs, rc, out.c_str());
  return rc < 0 ? 2 : 0;
}
#endif

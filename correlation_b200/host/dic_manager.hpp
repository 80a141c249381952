// dic_manager.hpp -- headless host for the GPU path: what managerClass does around CudaClass,
// without Qt / OpenCV. Restates, for `processor_GPU`:
//   perform_multiframe_correlation           manager_class.cpp:1297-1541 (frame loop, image rotation
//                                            :170-272, next-image prefetch :1438-1447)
//   perform_single_frame_correlation_*       :274-555 (rect), :557-814 (annulus), :1001-1295 (blob)
//   adjust_rectangular/annular/blob_domain   :2018-2310
//   adjust_initial_guess                     :2602-2707 (sector offset of the global guess on frame 0,
//                                            constant-velocity extrapolation for Eulerian + first-image)
//   update_results / update_global_results   :2312-2428, :2709-2753
//   initializeReport / addFrameToReport      :2473-2525, :2430-2471 (the CSV schema, verbatim columns)
// Frames are raw u8 buffers (the reference decodes files with cv::imread; decoding is out of scope).
#pragma once
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <future>
#include <sstream>
#include <string>
#include <vector>

#include "dic_cuda_class.hpp"

namespace dic_host {

enum domainEnum { domain_rectangular, domain_annular, domain_blob };

struct Config {
  domainEnum domain_type = domain_rectangular;
  // rectangularDomainStruct (domains.hpp:19-31)
  float x_begin = 0, y_begin = 0, x_end = 0, y_end = 0;
  int horizontal_subdivisions = 1, vertical_subdivisions = 1;
  // annularDomainStruct (domains.hpp:33-45)
  float x_center = 0, y_center = 0, r_inside = 0, r_outside = 0;
  int radial_subdivisions = 1, angular_subdivisions = 1;
  // blobDomainStruct (domains.hpp:47-57)
  v_points xy_contour;
  // solver (GUI defaults, mainapp.cpp:64,79,159,180,192-208)
  fittingModelEnum model = fm_UVUxUyVxVy;
  interpolationModelEnum interpolation = im_bicubic;
  int pyramid_start = 0, pyramid_step = 1, pyramid_stop = 2;
  float precision = 1e-3f;
  int max_iters = 50;
  deformationDescriptionEnum deformationDescription = def_Eulerian;
  referenceImageEnum referenceImage = refImage_First;
  errorHandlingModeEnum error_handling_mode = errorMode_stopAll; // GUI default, mainapp.cpp:73
  float global_initial_guess[DIC_MAX_PARAMS] = {0};
  int arith_mode = DIC_MODE_PARITY;
  bool batch_sectors = false; // extension: all sectors of a frame in one launch
  // extension: prefetch the next frame by ENQUEUEING its upload + pyramid on the engine's image stream instead of
  // running them on a per-frame loader thread (manager_class.cpp:1438-1447). Requires frame buffers that stay
  // alive and unchanged for the run (true for perform_multiframe_correlation's raw-buffer frames).
  // Measured on c3 (100 frames of 2048^2, B200): 2770 frames/s with the enqueue-only prefetch against 4170 with the
  // loader thread -- the pyramid kernels of frame k + 2, enqueued before frame k's solve, delay that solve's
  // cooperative launch (it needs all of its CTAs resident at once). Off by default; DIC_ASYNC_NEXT_IMAGE=1 turns it on.
  bool async_next_image = false;
};

// the part of frame_results (domains.hpp:59-108) the GPU path reads or reports
struct SectorState {
  float und_center_x = 0, und_center_y = 0, und_angle = 0;
  float und_global_center_x = 0, und_global_center_y = 0, und_global_angle = 0, und_global_ro = 0, und_global_ri = 0;
  float def_center_x = 0, def_center_y = 0, def_angle = 0, def_e = 0;
  float def_global_center_x = 0, def_global_center_y = 0, def_global_angle = 0;
  float past_und_center_x = 0, past_und_center_y = 0;
  float resulting_parameters[DIC_MAX_PARAMS] = {0};
  float previous_resulting_parameters[DIC_MAX_PARAMS] = {0};
  float initial_guess[DIC_MAX_PARAMS] = {0};
  int number_of_points = 0, iterations = 0;
  float chi = 0;
  bool error_status = false;
  errorEnum error_code = error_none;
};

inline int n_params_of(fittingModelEnum m) { return m == fm_U ? 1 : m == fm_UV ? 2 : m == fm_UVQ ? 3 : m == fm_UVUxUyVxVy ? 6 : 12; }

// interpolation_class.cpp:3-43 (+ quadratic extension): the model applied to one point
inline void distort_point(int model, float x, float y, float cx, float cy, const float *p, float &xd, float &yd) {
  float dx = x - cx, dy = y - cy;
  switch (model) {
  case fm_U: xd = x + p[0]; yd = y; break;
  case fm_UV: xd = x + p[0]; yd = y + p[1]; break;
  case fm_UVQ: xd = x + p[0] - dy * p[2]; yd = y + p[1] + dx * p[2]; break;
  case fm_UVUxUyVxVy: xd = x + p[0] + dx * p[2] + dy * p[3]; yd = y + p[1] + dx * p[4] + dy * p[5]; break;
  default:
    xd = x + p[0] + dx * p[2] + dy * p[3] + 0.5f * p[6] * dx * dx + p[7] * dx * dy + 0.5f * p[8] * dy * dy;
    yd = y + p[1] + dx * p[4] + dy * p[5] + 0.5f * p[9] * dx * dx + p[10] * dx * dy + 0.5f * p[11] * dy * dy;
  }
}

class HeadlessManager {
  Config cfg_;
  CudaClass cuda_;
  int np_;
  std::vector<SectorState> results_;
  std::ostringstream report_;
  // managerClass::error is ONE member overwritten by every sector's correlate (manager_class.cpp:443-452,
  // 704-713, 1145-1154): at the end of a frame it holds the status of the LAST sector processed
  bool error_ = false;
  int frames_done_ = 0;
  bool stops_frame() const { // :535-536, :795-796
    return error_ && (cfg_.error_handling_mode == errorMode_stopAll || cfg_.error_handling_mode == errorMode_stopFrame);
  }
  static constexpr float PI = 3.14159265359f; // parameters.hpp:23

  void distort(float x, float y, float cx, float cy, const float *p, float &xd, float &yd) const {
    distort_point(cfg_.model, x, y, cx, cy, p, xd, yd);
  }

  // manager_class.cpp:2527-2600 deformPoints: a contour (domain outline for plotting, def_contour of
  // :504 / :764 / :1207) carried along with a sector's result
  v_points deformPoints(const v_points &contour, float cx, float cy, const float *model_parameters) const {
    v_points out;
    out.reserve(contour.size());
    for (const auto &q : contour) {
      float xd, yd;
      distort(q.first, q.second, cx, cy, model_parameters, xd, yd);
      out.push_back(std::make_pair(xd, yd));
    }
    return out;
  }

  // manager_class.cpp:2602-2707
  void adjust_initial_guess(SectorState &s, int frame) {
    if (frame == 0) {
      for (int p = 0; p < np_; ++p) s.initial_guess[p] = cfg_.global_initial_guess[p];
      float dx = s.und_center_x - s.und_global_center_x, dy = s.und_center_y - s.und_global_center_y;
      const float *g = cfg_.global_initial_guess;
      if (cfg_.model == fm_UVQ) { // the reference applies this branch to U and UV too, reading g[2]; only UVQ has it
        s.initial_guess[0] += -dy * g[2];
        s.initial_guess[1] += dx * g[2];
      } else if (cfg_.model == fm_UVUxUyVxVy || cfg_.model == fm_UVUxUyVxVyQuad) {
        s.initial_guess[0] += dx * g[2] + dy * g[3];
        s.initial_guess[1] += dx * g[4] + dy * g[5];
      }
      for (int p = 0; p < np_; ++p) s.previous_resulting_parameters[p] = s.initial_guess[p];
    } else {
      if (cfg_.deformationDescription == def_Eulerian && cfg_.referenceImage == refImage_First) {
        for (int p = 0; p < np_; ++p) // constant rate of deformation
          s.initial_guess[p] = s.resulting_parameters[p] + (s.resulting_parameters[p] - s.previous_resulting_parameters[p]);
      } else {
        for (int p = 0; p < np_; ++p) s.initial_guess[p] = s.resulting_parameters[p];
      }
      for (int p = 0; p < np_; ++p) s.previous_resulting_parameters[p] = s.resulting_parameters[p];
    }
  }

  // frames other than the first: manager_class.cpp:2037-2084 (same for all three domain types)
  void carry_domain(SectorState &s) {
    if (cfg_.deformationDescription == def_Eulerian) return;
    s.und_global_center_x = s.def_global_center_x;
    s.und_global_center_y = s.def_global_center_y;
    s.und_global_angle = s.def_global_angle;
    s.past_und_center_x = s.und_center_x;
    s.past_und_center_y = s.und_center_y;
    s.und_center_x = s.def_center_x;
    s.und_center_y = s.def_center_y;
    s.und_angle = s.def_angle;
  }

  // manager_class.cpp:2312-2428 (GPU branch)
  void update_results(SectorState &s, const CorrelationResult &r) {
    s.chi = r.chi;
    s.number_of_points = r.numberOfPoints;
    s.iterations = r.iterations;
    s.error_code = r.errorCode;
    s.error_status = r.errorCode != error_none;
    s.und_center_x = r.undCenterX;
    s.und_center_y = r.undCenterY;
    for (int p = 0; p < np_; ++p) s.resulting_parameters[p] = r.resultingParameters[p];
    const float *p = s.resulting_parameters;
    switch (cfg_.model) {
    case fm_U: case fm_UV: s.def_angle = 0.f; break;
    case fm_UVQ: s.def_angle = p[2] + s.und_angle; break;
    default: // parameters.cpp:55-58
      s.def_angle = (float)std::atan2((double)(p[4] - p[3]), (double)(p[2] + p[5] + 2.f)) + s.und_angle;
    }
    s.def_e = 0.f;
    distort(s.und_center_x, s.und_center_y, s.und_center_x, s.und_center_y, p, s.def_center_x, s.def_center_y);
  }

  // manager_class.cpp:2709-2753
  void update_global_results() {
    float a = 0, cx = 0, cy = 0, tot = 0;
    for (auto &s : results_) {
      float n = (float)s.number_of_points;
      a += s.def_angle * n; cx += s.def_center_x * n; cy += s.def_center_y * n; tot += n;
    }
    a /= tot; cx /= tot; cy /= tot;
    for (auto &s : results_) { s.def_global_angle = a; s.def_global_center_x = cx; s.def_global_center_y = cy; }
  }

  void initializeReport() { // manager_class.cpp:2473-2525
    report_.str("");
    report_ << "Frame#,und_file_string,def_file_string,und_global_center_x,und_global_center_y,und_center_x,"
               "und_center_y,def_global_center_x,def_global_center_y,def_center_x,def_center_y,";
    for (int p = 0; p < np_; ++p) report_ << "parameter_" << p << ",";
    for (int p = 0; p < np_; ++p) report_ << "Initial_guess_" << p << ",";
    report_ << "und_global_angle(rad),def_global_angle(rad),und_angle(rad),def_angle(rad),def_angle(deg),"
               "chi,number_of_points,iterations,error_status,error_code" << std::endl;
  }
  void addFrameToReport(int frame, const std::string &und_name, const std::string &def_name) { // :2430-2471
    for (auto &s : results_) {
      report_ << frame << "," << und_name << "," << def_name << "," << s.und_global_center_x << ","
              << s.und_global_center_y << "," << s.und_center_x << "," << s.und_center_y << ","
              << s.def_global_center_x << "," << s.def_global_center_y << "," << s.def_center_x << ","
              << s.def_center_y << ",";
      for (int p = 0; p < np_; ++p) report_ << s.resulting_parameters[p] << ",";
      for (int p = 0; p < np_; ++p) report_ << s.initial_guess[p] << ",";
      report_ << s.und_global_angle << "," << s.def_global_angle << "," << s.und_angle << "," << s.def_angle << ","
              << s.def_angle * 180 / PI << ",";
      report_ << s.chi << "," << s.number_of_points << "," << s.iterations << "," << s.error_status << ","
              << s.error_code << std::endl;
    }
  }

  // ---- one frame, per domain type ------------------------------------------------------------
  bool frame_rectangular(int frame) {
    const int hs = cfg_.horizontal_subdivisions, vs = cfg_.vertical_subdivisions;
    const int x1 = (int)cfg_.x_end, x0 = (int)cfg_.x_begin, y1 = (int)cfg_.y_end, y0 = (int)cfg_.y_begin;
    const int xdim = (std::abs(x1 - x0) / hs - 1) / 2, ydim = (std::abs(y1 - y0) / vs - 1) / 2;
    const float fxdim = (std::fabs(cfg_.x_end - cfg_.x_begin) / (float)hs - 1.f) / 2.f;
    const float fydim = (std::fabs(cfg_.y_end - cfg_.y_begin) / (float)vs - 1.f) / 2.f;
    std::vector<float> guesses((size_t)hs * vs * np_);
    std::vector<int> boxes; // batch mode: every subset's rectangle, built in one call
    if (cfg_.batch_sectors && frame == 0) boxes.resize((size_t)hs * vs * 4);
    bool stop = false;
    for (int i = 0; i < hs && !stop; ++i) {
      int center_x = (int)(0.5f + cfg_.x_begin + fxdim + (2.f * fxdim + 1.f) * (float)i);
      for (int j = 0; j < vs; ++j) {
        const int iSector = i * vs + j;
        SectorState &s = results_[iSector];
        int center_y = (int)(0.5f + cfg_.y_begin + fydim + (2.f * fydim + 1.f) * (float)j);
        if (frame == 0) { // adjust_rectangular_domain :2021-2036
          s.und_global_center_x = (cfg_.x_begin + cfg_.x_end) * 0.5f;
          s.und_global_center_y = (cfg_.y_begin + cfg_.y_end) * 0.5f;
          s.und_global_angle = 0.f;
          s.und_center_x = (float)center_x; s.und_center_y = (float)center_y; s.und_angle = 0.f;
          s.past_und_center_x = s.und_center_x; s.past_und_center_y = s.und_center_y;
        } else {
          carry_domain(s);
        }
        center_x = (int)(s.und_center_x + 0.5f);
        center_y = (int)(s.und_center_y + 0.5f);
        adjust_initial_guess(s, frame);
        for (int p = 0; p < np_; ++p) guesses[(size_t)iSector * np_ + p] = s.initial_guess[p];
        if (cfg_.batch_sectors) {
          if (frame == 0) {
            int *b = &boxes[(size_t)iSector * 4];
            b[0] = center_x - xdim; b[1] = center_y - ydim; b[2] = center_x + xdim; b[3] = center_y + ydim;
          } else if (cfg_.deformationDescription != def_Eulerian && !s.error_status) {
            cuda_.updatePolygon(iSector, cfg_.deformationDescription);
          }
          continue;
        }
        if (frame == 0) {
          if (cuda_.resetPolygon(iSector, center_x - xdim, center_y - ydim, center_x + xdim, center_y + ydim) != error_none) {
            s.error_status = true; s.error_code = error_bad_domain; error_ = true;
            if (stops_frame()) { stop = true; break; }
            continue;
          }
        } else if (cfg_.deformationDescription != def_Eulerian) {
          cuda_.updatePolygon(iSector, cfg_.deformationDescription);
        }
        CorrelationResult *r = cuda_.correlate(iSector, &guesses[(size_t)iSector * np_]);
        update_results(s, *r);
        error_ = s.error_status; // :452
        if (stops_frame()) { stop = true; break; }
      }
    }
    if (cfg_.batch_sectors) {
      // extension: every subset of the frame in one launch. There is no "first sector that failed" inside a
      // launch, so stopFrame / stopAll act on the frame as a whole (error_ = any sector failed).
      const int n = hs * vs;
      bool bad_domain = false;
      if (frame == 0 && cuda_.resetPolygonGrid(0, n, boxes.data()) != error_none) bad_domain = true;
      std::vector<CorrelationResult> rs((size_t)n);
      const int rc = cuda_.correlateBatch(0, n, guesses.data(), rs.data());
      error_ = false;
      for (int k = 0; k < n; ++k) {
        if (rc == DIC_ERROR_BAD_ARGUMENT || rc == DIC_ERROR_CUDA) { // the launch itself failed: no record is valid
          results_[k].error_status = true; results_[k].error_code = (errorEnum)(rc == DIC_ERROR_CUDA ? error_cuda : error_bad_domain);
        } else {
          update_results(results_[k], rs[k]);
        }
        error_ = error_ || results_[k].error_status;
      }
      error_ = error_ || bad_domain;
    }
    update_global_results();
    return error_;
  }

  bool frame_annular(int frame) {
    const int rs = cfg_.radial_subdivisions, as = cfg_.angular_subdivisions;
    const float ri = cfg_.r_inside, ro = cfg_.r_outside;
    const float dr = (ro - ri) / (float)rs, da = 2.f * PI / (float)as;
    bool stop = false;
    for (int i = 0; i < rs && !stop; ++i)
      for (int j = 0; j < as; ++j) {
        const int iSector = i * as + j;
        SectorState &s = results_[iSector];
        if (frame == 0) { // adjust_annular_domain :2100-2142
          s.und_global_center_x = cfg_.x_center; s.und_global_center_y = cfg_.y_center;
          s.und_global_ro = ro; s.und_global_ri = ri; s.und_global_angle = 0.f;
          if (as > 1) {
            float center_angle = 0 + j * da + da / 2.f, center_r = ri + i * dr + dr / 2.f;
            s.und_center_x = s.und_global_center_x + center_r * (float)std::cos((double)center_angle);
            s.und_center_y = s.und_global_center_y + center_r * (float)std::sin((double)center_angle);
          } else {
            s.und_center_x = s.und_global_center_x; s.und_center_y = s.und_global_center_y;
          }
          s.past_und_center_x = s.und_center_x; s.past_und_center_y = s.und_center_y;
          s.und_angle = 0.f;
        } else {
          carry_domain(s);
        }
        const float cx = s.und_global_center_x, cy = s.und_global_center_y;
        const float r = ri + i * dr, a = s.und_global_angle + j * da;
        adjust_initial_guess(s, frame);
        if (frame == 0) {
          if (cuda_.resetPolygon(iSector, r, dr, a, da, cx, cy, as) != error_none) {
            s.error_status = true; s.error_code = error_bad_domain; error_ = true;
            if (stops_frame()) { stop = true; break; }
            continue;
          }
        } else if (cfg_.deformationDescription != def_Eulerian) {
          cuda_.updatePolygon(iSector, cfg_.deformationDescription);
        }
        float guess[DIC_MAX_PARAMS];
        for (int p = 0; p < np_; ++p) guess[p] = s.initial_guess[p];
        update_results(s, *cuda_.correlate(iSector, guess));
        error_ = s.error_status; // :713
        if (stops_frame()) { stop = true; break; } // :795-796
      }
    update_global_results();
    return error_;
  }

  bool frame_blob(int frame) {
    SectorState &s = results_[0];
    if (cfg_.xy_contour.size() < 3) { s.error_status = true; s.error_code = error_bad_domain; error_ = true; return true; }
    if (frame == 0) { // adjust_blob_domain :2247-2262, centre = mean of the contour (parameters.cpp:35-53)
      float xc = 0, yc = 0;
      for (auto &q : cfg_.xy_contour) { xc += q.first; yc += q.second; }
      xc /= (float)cfg_.xy_contour.size(); yc /= (float)cfg_.xy_contour.size();
      s.und_global_center_x = xc; s.und_global_center_y = yc; s.und_global_angle = 0.f;
      s.und_center_x = xc; s.und_center_y = yc; s.und_angle = 0.f;
      s.past_und_center_x = xc; s.past_und_center_y = yc;
    } else {
      carry_domain(s);
    }
    adjust_initial_guess(s, frame);
    if (frame == 0) {
      if (cuda_.resetPolygon(cfg_.xy_contour) != error_none) { // manager_class.cpp:1026-1030
        s.error_status = true; s.error_code = error_bad_domain; error_ = true; return true;
      }
    } else if (cfg_.deformationDescription != def_Eulerian) {
      cuda_.updatePolygon(0, cfg_.deformationDescription);
    }
    float guess[DIC_MAX_PARAMS];
    for (int p = 0; p < np_; ++p) guess[p] = s.initial_guess[p];
    update_results(s, *cuda_.correlate(0, guess));
    error_ = s.error_status; // :1154
    update_global_results();
    return error_;
  }

public:
  explicit HeadlessManager(const Config &cfg, int device = 0) : cfg_(cfg), cuda_(device), np_(n_params_of(cfg.model)) {
    if (cuda_.initialize() <= 0) throw std::runtime_error("no CUDA device: the GPU path has no CPU fallback");
    cuda_.set_max_iters(cfg.max_iters);           // mainapp.cpp:1677-1682
    cuda_.set_precision(cfg.precision);
    cuda_.set_fitting_model(cfg.model);
    cuda_.set_interpolation_model(cfg.interpolation);
    cuda_.set_arith_mode(cfg.arith_mode);
    if (const char *v = std::getenv("DIC_ASYNC_NEXT_IMAGE")) cfg_.async_next_image = v[0] == '1';
  }
  CudaClass &engine() { return cuda_; }
  int frames_done() const { return frames_done_; } // frame pairs that reached the report (stopAll ends early)
  const std::vector<SectorState> &results() const { return results_; }
  std::string report() const { return report_.str(); }

  // perform_multiframe_correlation. frames[k]: rows x cols u8. Returns the error flag.
  bool perform_multiframe_correlation(const std::vector<const uint8_t *> &frames, int rows, int cols,
                                      const std::vector<std::string> *names = nullptr) {
    const int n_sectors = cfg_.domain_type == domain_rectangular ? cfg_.horizontal_subdivisions * cfg_.vertical_subdivisions
                        : cfg_.domain_type == domain_annular ? cfg_.radial_subdivisions * cfg_.angular_subdivisions : 1;
    results_.assign(n_sectors, SectorState());
    frames_done_ = 0;
    initializeReport();
    if (frames.size() < 2) return true;
    auto name = [&](size_t k) { return names && k < names->size() ? (*names)[k] : ("frame" + std::to_string(k)); };
    // the GUI loads the first three images before the run (mainapp.cpp:900-912)
    cuda_.resetImagePyramids(frames[0], frames[1], frames.size() > 2 ? frames[2] : nullptr, rows, cols, color_monochrome,
                             cfg_.pyramid_start, cfg_.pyramid_step, cfg_.pyramid_stop);
    bool error = false;
    error_ = false;
    const int total_frame_pairs = (int)frames.size() - 1;
    // host-side phase times of the loop (DIC_HOST_PROFILE=1 prints them): rotation, prefetch start, the frame's
    // sectors (guess + correlate + result), wait for the prefetch, report row
    using clk = std::chrono::steady_clock;
    double t_phase[5] = {0, 0, 0, 0, 0};
    auto lap = [&](clk::time_point &t0, int k) { auto t1 = clk::now(); t_phase[k] += std::chrono::duration<double>(t1 - t0).count(); t0 = t1; };
    for (int frame = 0; frame < total_frame_pairs; ++frame) {
      clk::time_point tp = clk::now();
      if (frame > 0) {
        if (cfg_.referenceImage == refImage_Previous) cuda_.makeUndPyramidFromDef(); // :192-195
        cuda_.makeDefPyramidFromNxt();                                               // :233-235
      }
      lap(tp, 0);
      std::future<errorEnum> loader; // :1438-1447: next image upload + pyramid while this frame correlates
      const bool prefetch = frame + 2 < (int)frames.size() && frame > 0;
      errorEnum enqueue_rc = error_none;
      if (prefetch && cfg_.async_next_image)
        enqueue_rc = cuda_.resetNextPyramidAsync(frames[frame + 2], rows, cols); // stream-ordered, no thread
      else if (prefetch)
        loader = std::async(std::launch::async, [&, frame] { return cuda_.resetNextPyramid(frames[frame + 2], rows, cols); });
      lap(tp, 1);
      switch (cfg_.domain_type) {
      case domain_rectangular: error = frame_rectangular(frame); break;
      case domain_annular: error = frame_annular(frame); break;
      default: error = frame_blob(frame); break;
      }
      lap(tp, 2);
      if (prefetch && (cfg_.async_next_image ? enqueue_rc : loader.get()) != error_none) error = error_ = true; // :1469-1474 (error_multiThread)
      lap(tp, 3);
      addFrameToReport(frame, name(cfg_.referenceImage == refImage_First ? 0 : frame), name(frame + 1));
      frames_done_ = frame + 1;
      lap(tp, 4);
      if (error && cfg_.error_handling_mode == errorMode_stopAll) break; // :1493
    }
    if (std::getenv("DIC_HOST_PROFILE"))
      std::fprintf(stderr, "dic_host phases over %d frame pairs (ms): rotate %.3f, prefetch start %.3f, sectors %.3f, prefetch wait %.3f, report %.3f\n",
                   frames_done_, 1e3 * t_phase[0], 1e3 * t_phase[1], 1e3 * t_phase[2], 1e3 * t_phase[3], 1e3 * t_phase[4]);
    return error;
  }
};

} // namespace dic_host

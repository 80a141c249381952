// dic_cuda_class.hpp -- `CudaClass` over the C-ABI of libdic_b200.so.
//
// Same public methods, argument meaning and error behaviour as the reference facade
// (cuda_class.cuh:46-79) so that managerClass's call sites
// (manager_class.cpp:340, 449, 605, 710, 1039, 1152, 194, 234, 257) compile against it unchanged:
//   initialize / set_deviceCount / set_max_iters / set_precision / set_fitting_model /
//   set_interpolation_model / resetImagePyramids / resetNextPyramid / makeUndPyramidFromDef /
//   makeDefPyramidFromNxt / resetPolygon x3 / updatePolygon / correlate / getUndXY0ToCPU / getDefXY0ToCPU
// Image arguments: raw u8 buffers always; file paths only when DIC_WITH_OPENCV is defined (the
// reference decodes with cv::imread, cuda_class.cu:498-572 -- OpenCV is not part of this image).
#pragma once
#include <cstdint>
#include <cstring>
#include <deque>
#include <stdexcept>
#include <string>
#include <vector>

#include "dic_types.hpp"
#ifdef DIC_WITH_OPENCV
#include <opencv2/imgcodecs.hpp>
#endif

class CudaClass {
  dic_engine *engine_ = nullptr;
  int device_ = 0;
  int deviceCount = 0;
  int n_params_ = 6;
  // engine-owned result records, one per sector (cuda_polygon.cuh:370-371). A deque: growing it for a higher
  // sector id never moves the records already handed out by correlate()
  std::deque<CorrelationResult> results_;

  void need_engine() {
    if (!engine_) {
      engine_ = dic_create(device_);
      if (!engine_) throw std::runtime_error("dic_create failed: no CUDA device or out of memory");
    }
  }
  static int n_params_of(int m) { return m == 0 ? 1 : m == 1 ? 2 : m == 2 ? 3 : m == 3 ? 6 : 12; }

public:
  explicit CudaClass(int device = 0) : device_(device) {}
  ~CudaClass() { dic_destroy(engine_); }
  CudaClass(const CudaClass &) = delete;
  CudaClass &operator=(const CudaClass &) = delete;

  // cuda_class.cu:40-93: number of devices; 0 lets the caller fall back to its CPU engine
  int initialize() {
    deviceCount = dic_device_count();
    if (deviceCount > 0) need_engine();
    return deviceCount;
  }
  void set_deviceCount(int n) { deviceCount = n; }
  void set_max_iters(int n) { need_engine(); dic_set_max_iters(engine_, n); }
  void set_precision(float p) { need_engine(); dic_set_precision(engine_, p); }
  void set_fitting_model(fittingModelEnum m) { need_engine(); dic_set_fitting_model(engine_, (int)m); n_params_ = n_params_of((int)m); }
  void set_interpolation_model(interpolationModelEnum m) { need_engine(); dic_set_interpolation_model(engine_, (int)m); }
  // extensions
  void set_arith_mode(int mode) { need_engine(); dic_set_arith_mode(engine_, mode); }
  void set_center_mode(int mode) { need_engine(); dic_set_center_mode(engine_, mode); }
  dic_engine *handle() { need_engine(); return engine_; }

  // cuda_class.cu:512-572 (raw-buffer form)
  errorEnum resetImagePyramids(const uint8_t *und, const uint8_t *def, const uint8_t *nxt, int rows, int cols,
                               colorEnum color_mode, int start, int step, int stop) {
    need_engine();
    return (errorEnum)dic_reset_image_pyramids(engine_, und, def, nxt, rows, cols,
                                               color_mode == color_monochrome ? 1 : 3, start, step, stop);
  }
  errorEnum resetNextPyramid(const uint8_t *nxt, int rows, int cols) {
    need_engine();
    return (errorEnum)dic_reset_next_pyramid(engine_, nxt, rows, cols);
  }
  // extension: enqueue-only prefetch (no loader thread); nxt stays alive until makeDefPyramidFromNxt
  errorEnum resetNextPyramidAsync(const uint8_t *nxt, int rows, int cols) {
    need_engine();
    return (errorEnum)dic_reset_next_pyramid_async(engine_, nxt, rows, cols);
  }
#ifdef DIC_WITH_OPENCV
  void resetImagePyramids(const std::string undPath, const std::string defPath, const std::string nxtPath,
                          colorEnum color_mode, const int start, const int step, const int stop) {
    cv::Mat u = cv::imread(undPath, cv::IMREAD_GRAYSCALE), d = cv::imread(defPath, cv::IMREAD_GRAYSCALE);
    cv::Mat n = nxtPath.empty() ? cv::Mat() : cv::imread(nxtPath, cv::IMREAD_GRAYSCALE);
    resetImagePyramids(u.data, d.data, n.empty() ? nullptr : n.data, u.rows, u.cols, color_mode, start, step, stop);
  }
  void resetNextPyramid(const std::string nxtPath) {
    cv::Mat n = cv::imread(nxtPath, cv::IMREAD_GRAYSCALE);
    resetNextPyramid(n.data, n.rows, n.cols);
  }
#endif
  void makeUndPyramidFromDef() { need_engine(); dic_make_und_pyramid_from_def(engine_); }
  void makeDefPyramidFromNxt() { need_engine(); dic_make_def_pyramid_from_nxt(engine_); }

  // cuda_class.cu:574-605
  errorEnum resetPolygon(int iSector, int x0, int y0, int x1, int y1) {
    need_engine();
    return (errorEnum)dic_reset_polygon_rect(engine_, iSector, x0, y0, x1, y1);
  }
  errorEnum resetPolygon(int iSector, float r, float dr, float a, float da, float cx, float cy, int as) {
    need_engine();
    return (errorEnum)dic_reset_polygon_annular(engine_, iSector, r, dr, a, da, cx, cy, as);
  }
  errorEnum resetPolygon(v_points blobContour) { // the reference always uses sector 0 for blobs
    need_engine();
    std::vector<float> xy(2 * blobContour.size());
    for (size_t i = 0; i < blobContour.size(); ++i) { xy[2 * i] = blobContour[i].first; xy[2 * i + 1] = blobContour[i].second; }
    return (errorEnum)dic_reset_polygon_blob(engine_, 0, xy.data(), (int)blobContour.size());
  }
  void updatePolygon(int iSector, deformationDescriptionEnum d) { need_engine(); dic_update_polygon(engine_, iSector, (int)d); }

  // cuda_class.cu:104-293: guess is read, then overwritten with the result; the returned record
  // stays valid until the next correlate of the same sector. `results` is unused, as in the reference.
  template <class FrameResults>
  CorrelationResult *correlate(int iSector, float *initial_guess_, FrameResults & /*results*/) {
    return correlate(iSector, initial_guess_);
  }
  CorrelationResult *correlate(int iSector, float *initial_guess_) {
    need_engine();
    if ((int)results_.size() <= iSector) results_.resize(iSector + 1);
    dic_result r;
    std::memset(&r, 0, sizeof(r));
    int rc = dic_correlate(engine_, iSector, initial_guess_, &r);
    std::memcpy(&results_[iSector], &r, sizeof(r));
    if (rc >= DIC_ERROR_CUDA && r.errorCode == 0) results_[iSector].errorCode = (errorEnum)(rc > 7 ? error_cuda : rc);
    return &results_[iSector];
  }
  // extension: n rectangles at once (the subdivision loop's resetPolygon calls, manager_class.cpp:274-340)
  errorEnum resetPolygonGrid(int firstSector, int nSectors, const int *boxes) {
    need_engine();
    return (errorEnum)dic_reset_polygon_rect_grid(engine_, firstSector, nSectors, boxes);
  }
  // extension: all subsets of a subdivided domain in one launch
  int correlateBatch(int firstSector, int nSectors, float *guesses, CorrelationResult *out) {
    need_engine();
    return dic_correlate_batch(engine_, firstSector, nSectors, guesses, reinterpret_cast<dic_result *>(out));
  }

  // extension: the same as two calls -- the caller may stage the next pair or post-process the last records between them
  int correlateBatchAsync(int firstSector, int nSectors, const float *guesses) {
    need_engine();
    return dic_correlate_batch_async(engine_, firstSector, nSectors, guesses);
  }
  int correlateBatchWait(int firstSector, int nSectors, float *guessesOut, CorrelationResult *out) {
    need_engine();
    return dic_correlate_batch_wait(engine_, firstSector, nSectors, guessesOut, reinterpret_cast<dic_result *>(out));
  }
  // extension: whole (und, def) pairs from pinned host memory, up to two pairs ahead of the one being solved; the
  // loop `stage(k + 2); wait(k); advancePair(); correlateBatchAsync(k + 1)` keeps the PCIe bus busy all the time
  errorEnum stageNextPair(const unsigned char *und, const unsigned char *def, int rows, int cols) {
    need_engine();
    return (errorEnum)dic_stage_next_pair(engine_, und, def, rows, cols);
  }
  errorEnum advancePair() { need_engine(); return (errorEnum)dic_advance_pair(engine_); }

  v_points getUndXY0ToCPU(int iSector) { return points(iSector, false); }
  v_points getDefXY0ToCPU(int iSector) { return points(iSector, true); }

private:
  v_points points(int iSector, bool deformed) {
    need_engine();
    int64_t n = 0;
    auto fn = deformed ? dic_get_def_xy0 : dic_get_und_xy0;
    fn(engine_, iSector, nullptr, 0, &n);
    std::vector<float> xy(2 * (size_t)n + 2);
    fn(engine_, iSector, xy.data(), n, &n);
    v_points out((size_t)n);
    for (int64_t i = 0; i < n; ++i) out[i] = std::make_pair(xy[2 * i], xy[2 * i + 1]);
    return out;
  }
};

// oracle/shim/opencv2/core/core.hpp -- TEST INFRASTRUCTURE (oracle build shim).
// Minimal stand-in for the cv::Mat surface the reference CPU engine touches
// (pyramid_class.cpp:143-153 rows/cols/data/step1()/isContinuous();
//  correlation_class.hpp:56-58 value members and assignment). No arithmetic.
#ifndef ORACLE_SHIM_OPENCV_CORE_HPP
#define ORACLE_SHIM_OPENCV_CORE_HPP
// (the real core.hpp drags these in transitively; pyramid_class.cpp relies on it for assert/memcpy)
#include <cassert>
#include <cstddef>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#define CV_8U 0
#define CV_8UC1 0
#define CV_8UC3 16
namespace cv {
class Mat {
  std::shared_ptr<unsigned char> owner_;
  int channels_{1};

public:
  int rows{0};
  int cols{0};
  unsigned char *data{nullptr};
  Mat() {}
  // owning, zero-initialised
  Mat(int rows_in, int cols_in, int type)
      : channels_(type == CV_8UC3 ? 3 : 1), rows(rows_in), cols(cols_in) {
    size_t n = (size_t)rows * cols * channels_;
    owner_.reset(new unsigned char[n](), std::default_delete<unsigned char[]>());
    data = owner_.get();
  }
  // non-owning view on caller memory
  Mat(int rows_in, int cols_in, int type, void *ptr, size_t /*step*/ = 0)
      : channels_(type == CV_8UC3 ? 3 : 1), rows(rows_in), cols(cols_in),
        data((unsigned char *)ptr) {}
  size_t step1() const { return (size_t)cols * channels_; }
  bool isContinuous() const { return true; }
  bool empty() const { return data == nullptr || rows * cols == 0; }
  int channels() const { return channels_; }
};
inline void imshow(const char *, const Mat &) {}
} // namespace cv
#endif

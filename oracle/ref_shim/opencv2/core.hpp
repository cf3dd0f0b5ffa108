// TEST INFRASTRUCTURE — minimal stand-in for the parts of cv::Mat the reference's filter sources touch (see README.md).
#pragma once
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <memory>
#include <vector>

typedef unsigned char uchar;
#define CV_8UC1 0

namespace cv {

class Mat {
 public:
  int rows = 0, cols = 0;
  size_t step = 0;
  uchar* data = nullptr;

  Mat() {}
  Mat(int r, int c, int /*type*/) { create(r, c); }
  static Mat zeros(int r, int c, int type) { return Mat(r, c, type); }

  // unchecked element access, as OpenCV's release build
  template <typename T> T& at(int r, int c) { return *reinterpret_cast<T*>(data + (ptrdiff_t)r * (ptrdiff_t)step + (ptrdiff_t)c * (ptrdiff_t)sizeof(T)); }
  template <typename T> const T& at(int r, int c) const { return *reinterpret_cast<const T*>(data + (ptrdiff_t)r * (ptrdiff_t)step + (ptrdiff_t)c * (ptrdiff_t)sizeof(T)); }
  template <typename T> T& at(int i) { return *reinterpret_cast<T*>(data + (ptrdiff_t)i * (ptrdiff_t)sizeof(T)); }
  template <typename T> const T& at(int i) const { return *reinterpret_cast<const T*>(data + (ptrdiff_t)i * (ptrdiff_t)sizeof(T)); }

  Mat row(int r) const {
    Mat m;
    m.rows = 1; m.cols = cols; m.step = step; m.data = data + (size_t)r * step; m.store_ = store_;
    return m;
  }
  bool empty() const { return data == nullptr || rows == 0 || cols == 0; }

 private:
  static constexpr size_t kGuard = 64;  // zero bytes before and after the pixel buffer
  void create(int r, int c) {
    rows = r; cols = c; step = (size_t)c;
    store_ = std::make_shared<std::vector<uchar>>((size_t)r * (size_t)c + 2 * kGuard, (uchar)0);
    data = store_->data() + kGuard;
  }
  std::shared_ptr<std::vector<uchar>> store_;
};

}  // namespace cv

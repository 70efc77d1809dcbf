// Stand-in for <CImg.h>: the N x N float "information matrix" debug image of JointOptimization.cpp (info_mat_).
// Storage, element access, sizes; save_png() is a no-op (the image is a debug artefact outside the hot path).
#pragma once
#include <vector>
namespace cimg_library {
template <typename T>
struct CImg {
  int w, h;
  std::vector<T> px;
  CImg() : w(0), h(0) {}
  CImg(unsigned int w_, unsigned int h_, unsigned int = 1, unsigned int = 1, const T& v = T()) : w((int)w_), h((int)h_), px((size_t)w_ * h_, v) {}
  T& operator()(unsigned int x, unsigned int y) { return px[(size_t)y * w + x]; }
  const T& operator()(unsigned int x, unsigned int y) const { return px[(size_t)y * w + x]; }
  int width() const { return w; }
  int height() const { return h; }
  const CImg& save_png(const char*) const { return *this; }
};
}  // namespace cimg_library

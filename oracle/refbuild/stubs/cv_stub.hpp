/* Stand-in for the OpenCV header the reference rasteriser includes by absolute macOS
 * path (rasteriser/Source/skeleton.cpp:9).  OpenCV is not in this image, and one of the
 * texture files the reference loads (Textures/Marble2000x2000.jpg, :135) is not in its
 * repository, so the texture branches (skeleton.cpp:588-645) are exercised on SYNTHETIC
 * images: cv::Mat here is a plain byte image with OpenCV's addressing
 * (at<T>(row, col) = *(T *)(data + row * step + col * sizeof(T))), imread() returns a
 * deterministic procedural image keyed by the file name, and cvtColor / threshold do what
 * main() needs of them (:148-155).  The harnesses can also set the pixels directly.
 * Test infrastructure only. */
#ifndef B200_ORACLE_CV_STUB_HPP
#define B200_ORACLE_CV_STUB_HPP

#include <cstdint>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

typedef unsigned char uchar;

namespace cv {

struct Vec3b {
  uchar v[3];
  uchar operator[](int i) const { return v[i]; }
};

struct Mat {
  int rows = 0, cols = 0, channels = 0, step = 0;
  std::shared_ptr<std::vector<uchar>> store;   // shared on copy, like cv::Mat's reference count
  uchar *data = nullptr;
  void create(int r, int c, int ch) {
    rows = r; cols = c; channels = ch; step = c * ch;
    store = std::make_shared<std::vector<uchar>>((size_t)r * step);
    data = store->data();
  }
  template <typename T> T at(int r, int c) const {
    T t;
    std::memcpy(&t, data + (size_t)r * step + (size_t)c * sizeof(T), sizeof(T));
    return t;
  }
};

inline uint32_t stub_hash(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
  return x;
}

/* Procedural stand-in for a decoded texture: 2000 x 2000 for the marble, 1024 x 1024 for the rest
 * (the sizes findU / findV are called with, :591, :605), three channels.  "opacity" maps get a
 * lattice of dark holes so that both sides of the 100 threshold (:152-155) occur. */
inline Mat imread(const std::string &name, int) {
  Mat m;
  const int n = name.find("2000") != std::string::npos ? 2000 : 1024;
  m.create(n, n, 3);
  uint32_t seed = 2166136261u;
  for (char ch : name) seed = (seed ^ (uchar)ch) * 16777619u;
  const bool opacity = name.find("opacity") != std::string::npos;
  for (int r = 0; r < n; ++r)
    for (int c = 0; c < n; ++c) {
      uchar *p = m.data + (size_t)r * m.step + 3 * c;
      const uint32_t h = stub_hash(seed ^ (uint32_t)(r * n + c));
      if (opacity) {
        const bool hole = ((r >> 4) + (c >> 4)) % 3 == 0 || (h & 63u) == 0;
        const uchar v = hole ? (uchar)(h >> 8 & 63u) : (uchar)(160 + (h >> 8 & 63u));
        p[0] = p[1] = p[2] = v;
      } else {
        p[0] = (uchar)(64 + (h & 127u) + ((r >> 5) & 1) * 32);
        p[1] = (uchar)(32 + (h >> 8 & 127u) + ((c >> 5) & 1) * 64);
        p[2] = (uchar)(16 + (h >> 16 & 191u));
      }
    }
  return m;
}

/* CV_BGR2GRAY on 8-bit data: OpenCV's fixed-point weights (B 1868, G 9617, R 4899, >> 14). */
inline void cvtColor(const Mat &src, Mat &dst, int) {
  Mat out;
  out.create(src.rows, src.cols, 1);
  for (int r = 0; r < src.rows; ++r)
    for (int c = 0; c < src.cols; ++c) {
      const uchar *p = src.data + (size_t)r * src.step + (size_t)c * src.channels;
      const int b = p[0], g = src.channels > 1 ? p[1] : p[0], rr = src.channels > 2 ? p[2] : p[0];
      out.data[(size_t)r * out.step + c] = (uchar)((b * 1868 + g * 9617 + rr * 4899 + 8192) >> 14);
    }
  dst = out;
}

/* THRESH_BINARY (type 0): dst = src > thresh ? maxval : 0. */
inline double threshold(const Mat &src, Mat &dst, double thresh, double maxval, int) {
  Mat out;
  out.create(src.rows, src.cols, src.channels);
  for (size_t i = 0; i < (size_t)src.rows * src.step; ++i) out.data[i] = src.data[i] > thresh ? (uchar)maxval : (uchar)0;
  dst = out;
  return thresh;
}

}  // namespace cv

#define CV_LOAD_IMAGE_UNCHANGED (-1)
#define CV_BGR2GRAY 6

#endif

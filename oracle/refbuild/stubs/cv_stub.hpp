/* Inert stand-in for the OpenCV header the reference rasteriser includes by
 * absolute macOS path (rasteriser/Source/skeleton.cpp:9).  Only the texture
 * branches (texture != 0, out of scope) touch these types; the oracle always
 * runs with setting = settingBoxes = 0, so none of this is ever executed.
 * Test infrastructure only. */
#ifndef B200_ORACLE_CV_STUB_HPP
#define B200_ORACLE_CV_STUB_HPP

#include <string>
#include <vector>

typedef unsigned char uchar;

namespace cv {

struct Vec3b {
  uchar v[3];
  uchar operator[](int i) const { return v[i]; }
};

struct Mat {
  int rows = 0, cols = 0;
  template <typename T> T at(int, int) const { return T(); }
};

inline Mat imread(const std::string &, int) { return Mat(); }
inline void cvtColor(const Mat &, Mat &, int) {}
inline double threshold(const Mat &, Mat &, double, double, int) { return 0.0; }

}  // namespace cv

#define CV_LOAD_IMAGE_UNCHANGED (-1)
#define CV_BGR2GRAY 6

#endif

// cv_shim.h -- the handful of OpenCV types that appear in the reference's hot-path signatures
// (Main_Calibration/bundle_adjustment.h:52,63; bundle_adjustment_manager.h:11,14; reprojection_check.h:21),
// so that the source-compatible host classes build without OpenCV (absent from this image).  With real OpenCV
// available, compile with -DBA_HOST_USE_OPENCV and this header just includes <opencv2/core.hpp>.
#pragma once
#ifdef BA_HOST_USE_OPENCV
#include <opencv2/core.hpp>
#else
#include <cstddef>
#include <vector>

namespace cv {

struct Point2f { float x = 0.f, y = 0.f; Point2f() = default; Point2f(float x_, float y_) : x(x_), y(y_) {} };
struct Point2d { double x = 0., y = 0.; Point2d() = default; Point2d(double x_, double y_) : x(x_), y(y_) {} };
struct Point3d { double x = 0., y = 0., z = 0.; Point3d() = default; Point3d(double x_, double y_, double z_) : x(x_), y(y_), z(z_) {} };
struct Vec3d { double v[3] = {0., 0., 0.}; double& operator[](int i) { return v[i]; } const double& operator[](int i) const { return v[i]; } };

// dense row-major matrix of double (the only element type the path uses)
class Mat {
 public:
  int rows = 0, cols = 0;
  Mat() = default;
  Mat(int r, int c) : rows(r), cols(c), d_((size_t)r * c, 0.0) {}
  Mat(int r, int c, const double* src) : rows(r), cols(c), d_(src, src + (size_t)r * c) {}
  template <typename T> T& at(int r, int c = 0) { return d_[(size_t)r * cols + c]; }
  template <typename T> const T& at(int r, int c = 0) const { return d_[(size_t)r * cols + c]; }
  bool empty() const { return d_.empty(); }
  const double* ptr() const { return d_.data(); }
 private:
  std::vector<double> d_;
};

}  // namespace cv
#endif

// cv_storage.h -- just enough of cv::FileStorage's XML dialect for the files on either side of the path:
// Common/Calibration/Intrinsics/<serial>.xml (my_io.cpp:19-23) and Camera_Transform.xml
// (bundle_adjustment_manager.cpp:108-131, reprojection_check.cpp:35,65-66).  Dense double matrices only.
#pragma once
#include <cstdio>
#include <fstream>
#include <map>
#include <sstream>
#include <string>

#include "cv_shim.h"

namespace RSCalibration {
namespace storage {

inline bool ReadXml(const std::string& path, std::map<std::string, cv::Mat>& out) {
  std::ifstream f(path);
  if (!f) return false;
  std::stringstream ss;
  ss << f.rdbuf();
  const std::string s = ss.str();
  size_t pos = 0;
  const std::string tag = " type_id=\"opencv-matrix\">";
  while ((pos = s.find(tag, pos)) != std::string::npos) {
    const size_t lt = s.rfind('<', pos);
    const std::string name = s.substr(lt + 1, pos - lt - 1);
    auto field = [&](const char* k) -> std::string {
      const std::string open = std::string("<") + k + ">", close = std::string("</") + k + ">";
      const size_t a = s.find(open, pos), b = s.find(close, pos);
      return (a == std::string::npos || b == std::string::npos) ? std::string() : s.substr(a + open.size(), b - a - open.size());
    };
    const int rows = std::atoi(field("rows").c_str()), cols = std::atoi(field("cols").c_str());
    std::stringstream data(field("data"));
    cv::Mat m(rows, cols);
    for (int i = 0; i < rows * cols; ++i) data >> m.at<double>(i / cols, i % cols);
    out[name] = m;
    pos += tag.size();
  }
  return true;
}

// "%.17g" with OpenCV's trailing '.' for integral values
inline std::string Fmt17(double v) {
  char buf[64];
  std::snprintf(buf, sizeof(buf), "%.17g", v);
  std::string t(buf);
  const size_t e = t.find('e');
  std::string mant = e == std::string::npos ? t : t.substr(0, e);
  if (mant.find('.') == std::string::npos && mant.find("inf") == std::string::npos && mant.find("nan") == std::string::npos) mant += ".";
  if (e == std::string::npos) return mant;
  int ex = std::atoi(t.substr(e + 1).c_str());
  char eb[16];
  std::snprintf(eb, sizeof(eb), "e%+03d", ex);
  return mant + eb;
}

class XmlWriter {
 public:
  explicit XmlWriter(const std::string& path) : f_(path) { if (f_) f_ << "<?xml version=\"1.0\"?>\n<opencv_storage>\n"; }
  bool isOpened() const { return (bool)f_; }
  void Write(const std::string& name, const cv::Mat& m) {
    f_ << "<" << name << " type_id=\"opencv-matrix\">\n  <rows>" << m.rows << "</rows>\n  <cols>" << m.cols << "</cols>\n  <dt>d</dt>\n  <data>\n   ";
    size_t col = 3;
    for (int i = 0; i < m.rows * m.cols; ++i) {
      const std::string v = Fmt17(m.at<double>(i / m.cols, i % m.cols));
      if (col + 1 + v.size() > 72 && col > 3) { f_ << "\n   "; col = 3; }
      f_ << " " << v;
      col += 1 + v.size();
    }
    f_ << "</data></" << name << ">\n";
  }
  void release() { if (f_) { f_ << "</opencv_storage>\n"; f_.close(); } }
  ~XmlWriter() { release(); }
 private:
  std::ofstream f_;
};

}  // namespace storage
}  // namespace RSCalibration

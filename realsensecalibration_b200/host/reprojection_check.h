// reprojection_check.h -- RSCalibration::ReprojectionCheck::Reproject with the reference's signature
// (Main_Calibration/reprojection_check.h:16-22).  The numeric part (cv::projectPoints with zero distortion, the
// sum of ((x^ - x)^2 + (y^ - y)^2) / 2 over all corners and the per-coordinate RMS, reprojection_check.cpp:69,81,
// 100-101) runs on the GPU through ba_cuda_project_points_error_rt; drawing / imshow (reprojection_check.cpp:71-96)
// is GUI code outside the path and is not reproduced, so `images` is accepted and ignored.
#pragma once
#include <map>
#include <string>
#include <vector>

#include "cv_shim.h"
#include "my_const.h"

namespace RSCalibration {

class ReprojectionCheck {
 public:
  static void Reproject(std::vector<std::map<std::string, cv::Mat>>& images,
                        std::vector<std::vector<std::vector<cv::Point2f>>>& image_points_per_time,
                        std::map<std::string, cv::Mat>& camera_intrinsics_map, std::map<std::string, cv::Mat>& dist_coeffs_map);

  // additions: explicit paths / serial numbers (hard-coded in the reference, reprojection_check.cpp:7,35), results returned
  struct Result { double reprojection_error = 0.0, rms_per_coordinate = 0.0; long long num_points = 0; bool ok = false; };
  static Result Reproject(const std::string& point3d_path, const std::string& transform_xml_path,
                          const std::vector<std::string>& serial_numbers,
                          const std::vector<std::vector<std::vector<cv::Point2f>>>& image_points_per_time,
                          const std::map<std::string, cv::Mat>& camera_intrinsics_map, int device = 0);
};

}  // namespace RSCalibration

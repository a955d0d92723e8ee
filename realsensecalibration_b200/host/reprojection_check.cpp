#include "reprojection_check.h"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <iostream>

#include "ba_cuda.h"
#include "cv_storage.h"

namespace RSCalibration {

ReprojectionCheck::Result ReprojectionCheck::Reproject(const std::string& point3d_path, const std::string& transform_xml_path,
                                                       const std::vector<std::string>& serial_numbers,
                                                       const std::vector<std::vector<std::vector<cv::Point2f>>>& image_points_per_time,
                                                       const std::map<std::string, cv::Mat>& camera_intrinsics_map, int device) {
  Result res;
  FILE* fptr = std::fopen(point3d_path.c_str(), "r");
  if (fptr == NULL) { std::cerr << "unable to open point3d.txt" << std::endl; return res; }
  int num_points_all = 0, num_times = 0, num_cameras = 0;
  bool ok = std::fscanf(fptr, "%d", &num_points_all) == 1 && std::fscanf(fptr, "%d", &num_times) == 1 && std::fscanf(fptr, "%d", &num_cameras) == 1;
  std::vector<std::vector<int>> per(num_times, std::vector<int>(num_cameras, 0));
  for (int t = 0; ok && t < num_times; t++) {
    int tmp;
    ok = std::fscanf(fptr, "%d", &tmp) == 1;
    for (int c = 0; ok && c < num_cameras; c++) ok = std::fscanf(fptr, "%d", &per[t][c]) == 1;
  }
  std::map<std::string, cv::Mat> xf;
  if (!ok || !storage::ReadXml(transform_xml_path, xf)) {
    std::cerr << "unable to open Camera_Transform.xml" << std::endl;
    std::fclose(fptr);
    return res;
  }
  // the points follow in (time, camera, observation, corner) order (reprojection_check.cpp:47-62); the detected corner of
  // the same rank in image_points_per_time[time][camera] is its partner (:78)
  std::vector<double> xyz, rot(9 * (size_t)num_cameras), tv(3 * (size_t)num_cameras), intr(4 * (size_t)num_cameras);
  std::vector<int32_t> cam;
  std::vector<float> img;
  for (int t = 0; ok && t < num_times; t++)
    for (int c = 0; ok && c < num_cameras; c++)
      for (int i = 0; ok && i < per[t][c]; i++) {
        double p[3];
        ok = std::fscanf(fptr, "%lf", &p[0]) == 1 && std::fscanf(fptr, "%lf", &p[1]) == 1 && std::fscanf(fptr, "%lf", &p[2]) == 1 &&
             t < (int)image_points_per_time.size() && c < (int)image_points_per_time[t].size() && i < (int)image_points_per_time[t][c].size();
        if (!ok) break;
        xyz.insert(xyz.end(), p, p + 3);
        cam.push_back(c);
        img.push_back(image_points_per_time[t][c][i].x);
        img.push_back(image_points_per_time[t][c][i].y);
      }
  std::fclose(fptr);
  if (!ok) { std::cerr << "point3d.txt and the detected image points do not match" << std::endl; return res; }
  // R<i> is a 3x3 matrix in Main_Calibration's file (bundle_adjustment_manager.cpp:130) and the 3x1 rvec in Test2's
  // (Test2_BundleAdjustment/main.cpp:128); the rvec form goes to the entry point that runs cv::Rodrigues on the device
  std::vector<double> rvt(6 * (size_t)num_cameras);
  int n_mat = 0, n_rvec = 0;
  for (int c = 0; c < num_cameras; c++) {
    const cv::Mat& R = xf["R" + std::to_string(c)];
    const cv::Mat& t = xf["t" + std::to_string(c)];
    auto it = camera_intrinsics_map.find(serial_numbers[c]);
    if (R.empty() || t.empty() || it == camera_intrinsics_map.end()) { std::cerr << "missing transform / intrinsics of camera " << c << std::endl; return res; }
    if (R.rows * R.cols == 9) {
      for (int k = 0; k < 9; k++) rot[9 * (size_t)c + k] = R.at<double>(k / 3, k % 3);
      ++n_mat;
    } else if (R.rows * R.cols == 3) {
      for (int k = 0; k < 3; k++) rvt[6 * (size_t)c + k] = R.rows == 3 ? R.at<double>(k, 0) : R.at<double>(0, k);
      ++n_rvec;
    } else {
      std::cerr << "R" << c << " in " << transform_xml_path << " is neither a 3x3 matrix nor a 3x1 rotation vector" << std::endl;
      return res;
    }
    if (t.rows * t.cols != 3) { std::cerr << "t" << c << " in " << transform_xml_path << " is not a 3-vector" << std::endl; return res; }
    for (int k = 0; k < 3; k++) {
      tv[3 * (size_t)c + k] = t.rows == 3 ? t.at<double>(k, 0) : t.at<double>(0, k);
      rvt[6 * (size_t)c + 3 + k] = tv[3 * (size_t)c + k];
    }
    const cv::Mat& K = it->second;
    intr[4 * c + 0] = K.at<double>(0, 0); intr[4 * c + 1] = K.at<double>(1, 1); intr[4 * c + 2] = K.at<double>(0, 2); intr[4 * c + 3] = K.at<double>(1, 2);
  }
  if (n_mat != 0 && n_rvec != 0) { std::cerr << transform_xml_path << " mixes rotation matrices and rotation vectors" << std::endl; return res; }
  ba_cuda_problem* p = nullptr;
  double err = 0.0, rms = 0.0;
  const long long n = (long long)cam.size();
  ok = ba_cuda_create(&p, device) == BA_OK;
  if (ok && n_rvec == 0)
    ok = ba_cuda_project_points_error_rt(p, n, xyz.data(), cam.data(), num_cameras, rot.data(), tv.data(), intr.data(), img.data(), &err, nullptr, nullptr) == BA_OK;
  else if (ok)
    ok = ba_cuda_project_points_error(p, n, xyz.data(), cam.data(), num_cameras, rvt.data(), intr.data(), img.data(), &err, nullptr, nullptr) == BA_OK;
  if (!ok) std::cerr << "ReprojectionCheck: " << ba_cuda_last_error() << std::endl;
  ba_cuda_destroy(p);
  if (!ok) return res;
  rms = std::pow((err * 2.0) / (num_points_all * 2.0), 0.5);  // reprojection_check.cpp:101 divides by the header's count
  res.reprojection_error = err; res.rms_per_coordinate = rms; res.num_points = n; res.ok = true;
  return res;
}

void ReprojectionCheck::Reproject(std::vector<std::map<std::string, cv::Mat>>& images,
                                  std::vector<std::vector<std::vector<cv::Point2f>>>& image_points_per_time,
                                  std::map<std::string, cv::Mat>& camera_intrinsics_map, std::map<std::string, cv::Mat>& dist_coeffs_map) {
  (void)images; (void)dist_coeffs_map;  // overlays only / all-zero distortion in the reference's data
  const std::vector<std::string> serials(SERIAL_NUMBERS, SERIAL_NUMBERS + CAMERAS);
  const Result r = Reproject("../Common/Correspondence/hongo/point3d.txt", "../Common/Correspondence/hongo/Camera_Transform.xml", serials,
                             image_points_per_time, camera_intrinsics_map);
  if (!r.ok) std::exit(1);
  std::cout << "Reprojection Error (After BA): " << r.reprojection_error << std::endl;
  std::cout << "Average Reprojection Error per One Coordinate: " << r.rms_per_coordinate << std::endl;
}

}  // namespace RSCalibration

// test1_bundle_adjustmenter.h -- the Model A container of Test1_BundleAdjustment, source compatible with
// Test1_BundleAdjustment/bundle_adjustmenter.cpp:14-104 (global-namespace BALProblem there; namespaced here so
// that both containers can live in one binary), plus SolveTest1(): the replacement of that program's
// Problem / AddResidualBlock / Solve block (Test1_BundleAdjustment/main.cpp:65-87).
#pragma once
#include <cstdio>

#include "ba_cuda.h"
#include "cv_shim.h"

namespace RSCalibrationTest1 {

class BALProblem {
 public:
  BALProblem() = default;
  BALProblem(const BALProblem&) = delete;
  BALProblem& operator=(const BALProblem&) = delete;
  ~BALProblem() { delete[] point_index_; delete[] camera_index_; delete[] observations_; delete[] parameters_; }

  int num_observations() const { return num_observations_; }
  const double* observations() const { return observations_; }
  double* mutable_cameras() { return parameters_; }
  double* mutable_points() { return parameters_ + 6 * num_cameras_; }
  double* mutable_camera_for_observation(int i) { return mutable_cameras() + camera_index_[i] * 6; }
  double* mutable_point_for_observation(int i) { return mutable_points() + point_index_[i] * 3; }

  // "num_cameras num_points", then num_points x "cam pt x y" (num_observations == num_points, :64), then the
  // parameters.  A short read returns false (the reference LOG(FATAL)s, :88-95).
  bool LoadFile(const char* filename) {
    FILE* fptr = std::fopen(filename, "r");
    if (fptr == NULL) return false;
    bool ok = std::fscanf(fptr, "%d", &num_cameras_) == 1 && std::fscanf(fptr, "%d", &num_points_) == 1 && num_cameras_ > 0 && num_points_ >= 0;
    if (ok) {
      num_observations_ = num_points_;
      point_index_ = new int[num_observations_];
      camera_index_ = new int[num_observations_];
      observations_ = new double[2 * (size_t)num_observations_];
      num_parameters_ = 6 * num_cameras_ + 3 * num_points_;
      parameters_ = new double[num_parameters_];
      for (int i = 0; ok && i < num_observations_; ++i) {
        ok = std::fscanf(fptr, "%d", camera_index_ + i) == 1 && std::fscanf(fptr, "%d", point_index_ + i) == 1;
        for (int j = 0; ok && j < 2; ++j) ok = std::fscanf(fptr, "%lf", observations_ + 2 * (size_t)i + j) == 1;
      }
      for (int i = 0; ok && i < num_parameters_; ++i) ok = std::fscanf(fptr, "%lf", parameters_ + i) == 1;
    }
    std::fclose(fptr);
    return ok;
  }

  // additions for the C ABI
  int num_cameras() const { return num_cameras_; }
  int num_points() const { return num_points_; }
  int num_parameters() const { return num_parameters_; }
  const int* camera_index() const { return camera_index_; }
  const int* point_index() const { return point_index_; }
  double* mutable_parameters() { return parameters_; }

 private:
  int num_cameras_ = 0, num_points_ = 0, num_observations_ = 0, num_parameters_ = 0;
  int* point_index_ = nullptr;
  int* camera_index_ = nullptr;
  double* observations_ = nullptr;
  double* parameters_ = nullptr;
};

// One camera matrix for every observation, as Test1 does with serial_numbers[1] (main.cpp:73-74).
inline int SolveTest1(BALProblem& bal_problem, const cv::Mat& camera_matrix, ba_cuda_summary* summary, int device = 0,
                      bool progress_to_stdout = true) {
  const double intr[4] = {camera_matrix.at<double>(0, 0), camera_matrix.at<double>(1, 1), camera_matrix.at<double>(0, 2),
                          camera_matrix.at<double>(1, 2)};
  ba_cuda_problem* p = nullptr;
  int rc = ba_cuda_create(&p, device);
  if (rc == BA_OK)
    rc = ba_cuda_set_model_a(p, bal_problem.num_cameras(), bal_problem.num_points(), bal_problem.num_observations(),
                             bal_problem.camera_index(), bal_problem.point_index(), bal_problem.observations(), intr, 0);
  if (rc == BA_OK) rc = ba_cuda_set_parameters(p, bal_problem.mutable_parameters(), bal_problem.num_parameters());
  if (rc == BA_OK) {
    ba_cuda_options options;
    ba_cuda_options_init(&options);
    options.rcs_solver = BA_RCS_DENSE_CHOLESKY;
    options.minimizer_progress_to_stdout = progress_to_stdout ? 1 : 0;
    rc = ba_cuda_solve(p, &options, summary);
  }
  if (rc == BA_OK) rc = ba_cuda_get_parameters(p, bal_problem.mutable_parameters(), bal_problem.num_parameters());
  ba_cuda_destroy(p);
  return rc;
}

}  // namespace RSCalibrationTest1

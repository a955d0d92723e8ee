#include "bundle_adjustment.h"

#include <cstdio>
#include <cstring>

#include "ba_cuda.h"

namespace RSCalibration {

void BALProblem::release() {
  delete[] time_index_; delete[] camera_index_; delete[] marker_index_; delete[] observations_; delete[] parameters_;
  if (num_observations_per_time_camera_) {
    for (int i = 0; i < num_times_; i++) delete[] num_observations_per_time_camera_[i];
    delete[] num_observations_per_time_camera_;
  }
  time_index_ = camera_index_ = marker_index_ = nullptr;
  observations_ = parameters_ = nullptr;
  num_observations_per_time_camera_ = nullptr;
}

BALProblem::~BALProblem() { release(); }

int BALProblem::num_cameras() const { return num_cameras_; }
int BALProblem::num_observations() const { return num_observations_; }
// the file stores marker observations; callers want corner counts (bundle_adjustment.cpp:29-32)
int BALProblem::num_observations_per_time_camera(int time_idx, int camera_idx) const {
  return num_observations_per_time_camera_[time_idx][camera_idx] * 4;
}
const double* BALProblem::observations() const { return observations_; }
int BALProblem::num_parameters() const { return num_parameters_; }
const double* BALProblem::parameters() const { return parameters_; }
int BALProblem::num_times() const { return num_times_; }
int BALProblem::camera_idx(int observation_id) const { return camera_index_[observation_id]; }
int BALProblem::marker_idx(int observation_id) const { return marker_index_[observation_id]; }
double* BALProblem::camera_parameters(int camera_idx) { return parameters_ + 6 * camera_idx; }
double* BALProblem::marker_transform(int marker_idx) { return parameters_ + 6 * num_cameras_ + 6 * num_times_ + 6 * marker_idx; }
double* BALProblem::mutable_camera_transform_from_base_camera(int observation_idx) {
  return parameters_ + 6 * camera_index_[observation_idx];
}
double* BALProblem::mutable_base_marker_transform_from_base_camera(int observation_idx) {
  return parameters_ + 6 * num_cameras_ + 6 * time_index_[observation_idx];
}
double* BALProblem::mutable_marker_transform_from_base_marker(int observation_idx) {
  return parameters_ + 6 * num_cameras_ + 6 * num_times_ + 6 * marker_index_[observation_idx];
}

// Same text format and read order as the reference (bundle_adjustment.cpp:132-187); unlike it, short reads are
// reported (the reference ignores fscanf's return value) and the file is closed.
bool BALProblem::loadFile(const char* filename) {
  FILE* fptr = std::fopen(filename, "r");
  if (fptr == NULL) return false;
  release();
  bool ok = std::fscanf(fptr, "%d", &num_times_) == 1 && std::fscanf(fptr, "%d", &num_cameras_) == 1 &&
            std::fscanf(fptr, "%d", &num_markers_) == 1 && std::fscanf(fptr, "%d", &num_observations_) == 1 &&
            num_times_ >= 0 && num_cameras_ > 0 && num_markers_ > 0 && num_observations_ >= 0;
  if (!ok) { std::fclose(fptr); num_times_ = num_cameras_ = num_markers_ = num_observations_ = 0; return false; }
  time_index_ = new int[num_observations_];
  camera_index_ = new int[num_observations_];
  marker_index_ = new int[num_observations_];
  observations_ = new double[8 * (size_t)num_observations_];
  num_observations_per_time_camera_ = new int*[num_times_];
  for (int i = 0; i < num_times_; i++) num_observations_per_time_camera_[i] = new int[num_cameras_];
  num_parameters_ = 6 * num_cameras_ + 6 * num_times_ + 6 * num_markers_;
  parameters_ = new double[num_parameters_];
  for (int time_idx = 0; ok && time_idx < num_times_; time_idx++) {
    int tmp;
    ok = std::fscanf(fptr, "%d", &tmp) == 1;
    for (int camera_idx = 0; ok && camera_idx < num_cameras_; camera_idx++)
      ok = std::fscanf(fptr, "%d", &num_observations_per_time_camera_[time_idx][camera_idx]) == 1;
  }
  for (int i = 0; ok && i < num_observations_; i++) {
    ok = std::fscanf(fptr, "%d", time_index_ + i) == 1 && std::fscanf(fptr, "%d", camera_index_ + i) == 1 &&
         std::fscanf(fptr, "%d", marker_index_ + i) == 1;
    for (int j = 0; ok && j < 8; j++) ok = std::fscanf(fptr, "%lf", observations_ + 8 * (size_t)i + j) == 1;
  }
  for (int i = 0; ok && i < num_parameters_; i++) ok = std::fscanf(fptr, "%lf", parameters_ + i) == 1;
  std::fclose(fptr);
  return ok;
}

// marker corner -> base-marker frame -> base camera for every observation x 4 corners
// (bundle_adjustment.cpp:89-130); composed on the GPU by ba_cuda_model_b_outputs.
void BALProblem::getPoint3dCoordinates(std::vector<cv::Point3d>& points) {
  if (num_observations_ == 0) return;
  ba_cuda_problem* p = nullptr;
  std::vector<double> corners(12 * (size_t)num_observations_), intr(4 * (size_t)num_cameras_, 1.0);
  bool ok = ba_cuda_create(&p, 0) == BA_OK &&
            ba_cuda_set_model_b(p, num_cameras_, num_times_, num_markers_, num_observations_, time_index_, camera_index_, marker_index_,
                                observations_, intr.data(), marker_side_, 1, 1) == BA_OK &&
            ba_cuda_set_parameters(p, parameters_, num_parameters_) == BA_OK &&
            ba_cuda_model_b_outputs(p, nullptr, nullptr, corners.data()) == BA_OK;
  if (!ok) std::fprintf(stderr, "getPoint3dCoordinates: %s\n", ba_cuda_last_error());
  ba_cuda_destroy(p);
  if (!ok) return;
  for (size_t i = 0; i < corners.size(); i += 3) points.emplace_back(cv::Point3d(corners[i], corners[i + 1], corners[i + 2]));
}

}  // namespace RSCalibration

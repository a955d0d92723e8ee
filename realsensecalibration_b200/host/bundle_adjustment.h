// bundle_adjustment.h -- RSCalibration::BALProblem, source compatible with the reference's container
// (Main_Calibration/bundle_adjustment.h:18-54, bundle_adjustment.cpp:5-187): same public methods, same flat
// parameters_ = [cameras | frames | markers] x 6 and observations_ = 8 per marker observation, same file
// format.  The four AutoDiff functors that followed it in the reference header (bundle_adjustment.h:56-343) have
// no host-side counterpart any more: their arithmetic lives in the CUDA kernels behind ba_cuda_set_model_b().
#pragma once
#include <string>
#include <vector>

#include "cv_shim.h"
#include "my_const.h"

namespace RSCalibration {

class BALProblem {
 private:
  int num_times_ = 0, num_cameras_ = 0, num_markers_ = 0, num_observations_ = 0, num_parameters_ = 0;
  int* time_index_ = nullptr;
  int* camera_index_ = nullptr;
  int* marker_index_ = nullptr;
  int** num_observations_per_time_camera_ = nullptr;
  double* observations_ = nullptr;
  double* parameters_ = nullptr;
  double marker_side_ = MARKER_SIDE;
  void release();

 public:
  BALProblem() = default;
  BALProblem(const BALProblem&) = delete;
  BALProblem& operator=(const BALProblem&) = delete;
  ~BALProblem();

  int num_cameras() const;
  int num_observations() const;
  int num_observations_per_time_camera(int time_idx, int camera_idx) const;
  const double* observations() const;
  int num_parameters() const;
  const double* parameters() const;
  int num_times() const;
  int camera_idx(int observation_id) const;
  int marker_idx(int observation_id) const;
  double* camera_parameters(int camera_idx);
  double* marker_transform(int marker_idx);
  double* mutable_camera_transform_from_base_camera(int observation_idx);
  double* mutable_base_marker_transform_from_base_camera(int observation_idx);
  double* mutable_marker_transform_from_base_marker(int observation_idx);
  void getPoint3dCoordinates(std::vector<cv::Point3d>& points);
  bool loadFile(const char* filename);

  // additions (not in the reference): what the C ABI needs as plain arrays, and the run-time marker side
  int num_markers() const { return num_markers_; }
  const int* time_index() const { return time_index_; }
  const int* camera_index() const { return camera_index_; }
  const int* marker_index() const { return marker_index_; }
  double* mutable_parameters() { return parameters_; }
  void set_marker_side(double side) { marker_side_ = side; }
  double marker_side() const { return marker_side_; }
};

}  // namespace RSCalibration

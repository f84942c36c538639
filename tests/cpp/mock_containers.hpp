// Minimal stand-ins for the reference containers (covins::Keyframe / Landmark / Map / LoopConstraint) exposing exactly
// the member names covins_b200_shim.hpp uses.  They let the shim be compiled and exercised without Eigen/OpenCV/aslam/
// robopt (absent offline).  Field names follow keyframe_base.hpp:159-237, landmark_base.hpp:87-119, map_base.hpp:97-112.
#pragma once
#include <array>
#include <cstddef>
#include <map>
#include <memory>
#include <utility>
#include <vector>

namespace mock {

struct Transform {   // Eigen::Matrix4d stand-in
  double m[16];
  static Transform Identity() {
    Transform T{};
    for (int i = 0; i < 16; i++) T.m[i] = (i % 5 == 0) ? 1.0 : 0.0;
    return T;
  }
  double& operator()(int r, int c) { return m[4 * r + c]; }
  double operator()(int r, int c) const { return m[4 * r + c]; }
};
struct Matrix6 {
  double m[36];
  double operator()(int r, int c) const { return m[6 * r + c]; }
};
using Vector3 = std::array<double, 3>;
using Vector2f = std::array<float, 2>;
using Vector4f = std::array<float, 4>;
using idpair = std::pair<size_t, size_t>;

struct Landmark;
struct Keyframe;
using KeyframePtr = std::shared_ptr<Keyframe>;
using LandmarkPtr = std::shared_ptr<Landmark>;

struct Keyframe : std::enable_shared_from_this<Keyframe> {
  idpair id_;
  bool invalid = false, is_loaded_ = false, is_gba_optimized_ = false;
  Transform T_w_s_ = Transform::Identity(), T_w_s_vio_ = Transform::Identity(), T_s_c_ = Transform::Identity();
  Vector3 velocity_{0, 0, 0}, bias_accel_{0, 0, 0}, bias_gyro_{0, 0, 0};
  std::vector<Vector2f> keypoints_distorted_;
  std::vector<Vector4f> keypoints_aors_;   // angle, octave, response, size
  std::vector<LandmarkPtr> landmarks_;
  KeyframePtr pred, succ;
  double intr[4], dist[4];
  std::vector<double> imu_dt, imu_acc, imu_gyr;
  double imu_acc0[3] = {0, 0, 0}, imu_gyr0[3] = {0, 0, 0}, imu_noise[5] = {0, 0, 0, 0, 9.81};
  int n_pose_optimized = 0, n_velbias_optimized = 0;
  // what FeatureMatcher::SearchBySE3 reads (feature_matcher_be.cpp:293-498)
  std::vector<std::array<unsigned char, 32>> descriptors_;   // cv::Mat descriptors_ rows
  double K_[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  Transform T_c_w_ = Transform::Identity();
  double img_dim_x_min_ = 0, img_dim_x_max_ = 752, img_dim_y_min_ = 0, img_dim_y_max_ = 480;
  const unsigned char* GetDescriptor(size_t i) const { return descriptors_[i].data(); }   // keyframe_base.cpp:254-256
  std::vector<LandmarkPtr> GetLandmarks() const { return landmarks_; }
  LandmarkPtr GetLandmark(size_t i) const { return landmarks_[i]; }
  void RemapLandmark(LandmarkPtr lm, size_t feat_id_now, size_t feat_id_new);       // keyframe_be.cpp:484-495, defined below Landmark
  Transform GetPoseTcw() const { return T_c_w_; }
  double calibration_K(int r, int c) const { return K_[3 * r + c]; }                       // calibration_.K(r, c)
  double image_width() const { return 752.0; }                                              // camera_->imageWidth()
  double image_height() const { return 480.0; }

  bool IsInvalid() const { return invalid; }
  Transform GetPoseTws() const { return T_w_s_; }
  Transform GetStateExtrinsics() const { return T_s_c_; }
  Transform GetPoseTws_vio() const { return T_w_s_vio_; }
  void SetPoseTws(const Transform& T) { T_w_s_ = T; }
  Vector3 GetStateVelocity() const { return velocity_; }
  void SetStateVelocity(const Vector3& v) { velocity_ = v; }
  void SetStateBias(const Vector3& ba, const Vector3& bg) { bias_accel_ = ba; bias_gyro_ = bg; }
  void SetPoseOptimized() { n_pose_optimized++; }
  void SetVelBiasOptimized() { n_velbias_optimized++; }
  KeyframePtr GetPredecessor() const { return pred; }
  KeyframePtr GetSuccessor() const { return succ; }
  void EraseLandmark(size_t kp) { landmarks_[kp] = nullptr; }
  // KeyframeBase::UpdateCeresFromState (keyframe_base.cpp:486-521)
  void UpdateCeresFromState(double* pose, double* vb, double* extr) const;
  // what Adapter<> needs
  bool GetCameraParams(double i[4], double d[4]) const {
    for (int k = 0; k < 4; k++) { i[k] = intr[k]; d[k] = dist[k]; }
    return true;
  }
  const std::vector<double>& ImuDt() const { return imu_dt; }
  void GetImu(std::vector<double>& dt, std::vector<double>& acc, std::vector<double>& gyr, double a0[3], double g0[3],
              double noise[5]) const {
    dt = imu_dt; acc = imu_acc; gyr = imu_gyr;
    for (int k = 0; k < 3; k++) { a0[k] = imu_acc0[k]; g0[k] = imu_gyr0[k]; }
    for (int k = 0; k < 5; k++) noise[k] = imu_noise[k];
  }
};

struct Landmark {
  idpair id_;
  bool invalid = false, is_gba_optimized_ = false;
  Vector3 pos_w_{0, 0, 0};
  std::map<KeyframePtr, size_t> observations_;   // pointer-ordered like the reference (typedefs_base.hpp:187)
  KeyframePtr ref_kf;
  int n_optimized = 0;
  double max_distance_ = 1.0, min_distance_ = 0.0;
  Vector3 normal_{0, 0, 0};
  std::array<unsigned char, 32> descriptor_{};
  double GetMaxDistanceInvariance() const { return 1.2 * max_distance_; }                    // landmark_base.cpp:68-71
  double GetMaxDistance() const { return max_distance_; }                                   // the added getter (PredictScale's operand)
  double GetMinDistanceInvariance() const { return 0.8 * min_distance_; }                    // landmark_base.cpp:73-76
  Vector3 GetNormal() const { return normal_; }
  void AddObservation(const KeyframePtr& kf, size_t idx) { observations_[kf] = idx; }      // (the plain map insert of LandmarkBase::AddObservation)
  const unsigned char* GetDescriptorPtr() const { return descriptor_.data(); }              // Landmark::GetDescriptor().data
  int GetFeatureIndex(const KeyframePtr& kf) const {
    auto it = observations_.find(kf);
    return it == observations_.end() ? -1 : (int)it->second;
  }
  bool IsInvalid() const { return invalid; }
  std::map<KeyframePtr, size_t> GetObservations() const { return observations_; }
  Vector3 GetWorldPos() const { return pos_w_; }
  void SetWorldPos(const Vector3& p) { pos_w_ = p; }
  void SetOptimized() { n_optimized++; }
  void EraseObservation(const KeyframePtr& kf) { observations_.erase(kf); }
  KeyframePtr GetReferenceKeyframe() const { return ref_kf; }
};

inline void Keyframe::RemapLandmark(LandmarkPtr lm, size_t feat_id_now, size_t feat_id_new) {   // statement order of the reference
  auto lm_new = landmarks_[feat_id_new];
  landmarks_[feat_id_now] = nullptr;
  landmarks_[feat_id_new] = lm;
  lm->EraseObservation(shared_from_this());
  lm->AddObservation(shared_from_this(), feat_id_new);
  if (lm_new) lm_new->EraseObservation(shared_from_this());
}

struct LoopConstraint {   // typedefs_base.hpp:264-277
  KeyframePtr kf1, kf2;
  Transform T_s1_s2;
  Matrix6 cov_mat;
};

struct Map {
  size_t id_map_ = 0;
  std::map<idpair, KeyframePtr> keyframes_;
  std::map<idpair, LandmarkPtr> landmarks_;
  std::vector<LoopConstraint> loops_;
  int n_clean = 0;
  std::vector<KeyframePtr> GetKeyframesVec() const {
    std::vector<KeyframePtr> v;
    for (auto& p : keyframes_) v.push_back(p.second);
    return v;
  }
  std::vector<LandmarkPtr> GetLandmarksVec() const {
    std::vector<LandmarkPtr> v;
    for (auto& p : landmarks_) v.push_back(p.second);
    return v;
  }
  std::vector<LoopConstraint> GetLoopConstraints() const { return loops_; }
  void EraseLandmark(const LandmarkPtr& lm) { landmarks_.erase(lm->id_); }
  void Clean() { n_clean++; }
};

}  // namespace mock

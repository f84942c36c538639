// covins_b200_shim.hpp — host-side C++ shim that keeps the reference's own call surface and routes it to the
// C-ABI of libcovins_b200.so (include/covins_b200.h).  Header-only, C++17, no third-party headers.
//
// The reference has no plugin/FFI layer; its seams are C++ symbols:
//   Optimization::GlobalBundleAdjustment(MapPtr, int, double, bool, bool, bool)   optimization_be.hpp:38-40
//   Optimization::PoseGraphOptimization(MapPtr, PoseMap)                          optimization_be.hpp:46-47
//   the per-candidate matching block of PlaceRecognitionG::ComputeSE3             placerec_gen_be.cpp:72-125
//   the per-candidate matching block of PlaceRecognition::ComputeSE3              placerec_be.cpp:75-113
// Each function below has the same name, argument meaning and write-back behaviour; the body is
//   flatten containers (canonical orders) → one C-ABI call → scatter through the reference's setters
// in the same order the reference calls them.
//
// The functions are templates over the container types, and touch them ONLY through member names the reference
// classes already have (KeyframeBase/Keyframe: keyframe_base.hpp:159-237, keyframe_be.hpp:86-112; LandmarkBase/
// Landmark: landmark_base.hpp:87-119, landmark_be.hpp:60-77; MapBase/Map: map_base.hpp:97-112; LoopConstraint:
// typedefs_base.hpp:264-277).  Inside the covins_backend tree they instantiate with the real classes (INTEGRATION.md
// shows the two-line change in optimization_be.cpp / placerec_gen_be.cpp); in this repository they are instantiated
// with the mock containers of tests/cpp/mock_containers.hpp.  Three things the real classes reach through
// third-party types are funnelled through one adapter, covins_b200::Adapter<Keyframe>, which the integrator
// specialises (camera intrinsics / distortion from aslam::Camera, raw IMU samples from robopt PreintegrationBase).
#pragma once
#include <algorithm>
#include <array>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <map>
#include <memory>
#include <set>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "../../../include/covins_b200.h"

namespace covins_b200 {

// ---------------------------------------------------------------------------------------------------------------
// RAII context; one per host thread (the reference runs one place-recognition thread per agent).
// ---------------------------------------------------------------------------------------------------------------
class Context {
 public:
  explicit Context(int device = 0) {
    if (cvb_ctx_create(device, &ctx_) != CVB_OK)
      throw std::runtime_error("covins_b200: no usable CUDA device (there is no CPU fallback)");
  }
  ~Context() { cvb_ctx_destroy(ctx_); }
  Context(const Context&) = delete;
  Context& operator=(const Context&) = delete;
  cvb_ctx* get() const { return ctx_; }
  // error convention of the reference: print + exit(-1) (optimization_be.cpp:113-114); soft failures return
  void check(int rc, const char* what) const {
    if (rc == CVB_OK) return;
    std::fprintf(stderr, "\033[1;31m!!!!! FATAL !!!!!\033[0m covins_b200 %s: status %d: %s\n", what, rc, cvb_last_error(ctx_));
    std::exit(-1);
  }

 private:
  cvb_ctx* ctx_ = nullptr;
};

// ---------------------------------------------------------------------------------------------------------------
// Adapter: the only place that touches third-party types of the real containers.  Default = mock containers.
// ---------------------------------------------------------------------------------------------------------------
template <class KF>
struct Adapter {
  // camera: pinhole intrinsics [fx,fy,cx,cy] + radtan [k1,k2,p1,p2]   (aslam::PinholeCamera::getParameters(),
  // getDistortion().getParameters(); optimization_be.cpp:95-103).  Returns false for unsupported models.
  static bool camera(const KF& kf, double intr[4], double dist[4]) { return kf.GetCameraParams(intr, dist); }
  // camera / distortion type of the keyframe's aslam camera — the template arguments the reference picks for
  // GlobalEuclideanReprError at optimization_be.cpp:186-231: cam 0 = kPinhole, 1 = kUnifiedProjection (xi = its first
  // intrinsic); dist 0 = kRadTan, 1 = kEquidistant, 2 = kFisheye.  The default serves containers without the notion
  // (ORB-SLAM3 agents only send pinhole + radtan, orb_slam3/src/KeyFrame.cc:64-65); specialise for the real Keyframe:
  //   cam = kf.camera_->getType() == aslam::Camera::Type::kUnifiedProjection, dist from getDistortion().getType().
  static void camera_model(const KF& kf, int* cam, int* dist, double* xi) { (void)kf; *cam = 0; *dist = 0; *xi = 0.0; }
  // raw IMU samples of the KF's preintegration (robopt PreintegrationBase::getReadingsByIndex / getTimeDiffByIndex,
  // keyframe_base.cpp:145-173) and its first reading + noise (keyframe_be.cpp:187-203)
  static size_t imu_count(const KF& kf) { return kf.ImuDt().size(); }
  static void imu_samples(const KF& kf, std::vector<double>& dt, std::vector<double>& acc, std::vector<double>& gyr,
                          double acc0[3], double gyr0[3], double noise[5]) {
    kf.GetImu(dt, acc, gyr, acc0, gyr0, noise);
  }
};

namespace detail {

// rotation matrix (via operator()(r,c) of a 4x4 transform) → quaternion (x,y,z,w), the convention of
// Eigen::Quaterniond(R) used at keyframe_base.cpp:490-499 (w >= 0 branch of Eigen's algorithm)
template <class T4>
inline void transform_to_pose7(const T4& T, double* out) {
  const double m00 = T(0, 0), m11 = T(1, 1), m22 = T(2, 2);
  double q[4];  // x y z w
  const double tr = m00 + m11 + m22;
  if (tr > 0) {
    double t = std::sqrt(tr + 1.0);
    q[3] = 0.5 * t;
    t = 0.5 / t;
    q[0] = (T(2, 1) - T(1, 2)) * t;
    q[1] = (T(0, 2) - T(2, 0)) * t;
    q[2] = (T(1, 0) - T(0, 1)) * t;
  } else {
    int i = 0;
    if (m11 > m00) i = 1;
    if (m22 > T(i, i)) i = 2;
    const int j = (i + 1) % 3, k = (j + 1) % 3;
    double t = std::sqrt(T(i, i) - T(j, j) - T(k, k) + 1.0);
    q[i] = 0.5 * t;
    t = 0.5 / t;
    q[3] = (T(k, j) - T(j, k)) * t;
    q[j] = (T(j, i) + T(i, j)) * t;
    q[k] = (T(k, i) + T(i, k)) * t;
  }
  out[0] = q[0]; out[1] = q[1]; out[2] = q[2]; out[3] = q[3];
  out[4] = T(0, 3); out[5] = T(1, 3); out[6] = T(2, 3);
}

// Utils::Ceres2Transform (covins_comm/src/covins_base/utils_base.cpp:28-43): normalised quaternion → 4x4
template <class T4>
inline T4 pose7_to_transform(const double* p) {
  double x = p[0], y = p[1], z = p[2], w = p[3];
  const double n = 1.0 / std::sqrt(x * x + y * y + z * z + w * w);
  x *= n; y *= n; z *= n; w *= n;
  T4 T = T4::Identity();
  T(0, 0) = 1 - 2 * (y * y + z * z); T(0, 1) = 2 * (x * y - z * w); T(0, 2) = 2 * (x * z + y * w);
  T(1, 0) = 2 * (x * y + z * w); T(1, 1) = 1 - 2 * (x * x + z * z); T(1, 2) = 2 * (y * z - x * w);
  T(2, 0) = 2 * (x * z - y * w); T(2, 1) = 2 * (y * z + x * w); T(2, 2) = 1 - 2 * (x * x + y * y);
  T(0, 3) = p[4]; T(1, 3) = p[5]; T(2, 3) = p[6];
  return T;
}

template <class T4>
inline T4 rel_transform(const T4& Ta, const T4& Tb) {  // Ta^-1 * Tb for rigid transforms
  T4 R = T4::Identity();
  for (int r = 0; r < 3; r++)
    for (int c = 0; c < 3; c++) {
      double s = 0;
      for (int m = 0; m < 3; m++) s += Ta(m, r) * Tb(m, c);
      R(r, c) = s;
    }
  for (int r = 0; r < 3; r++) {
    double s = 0;
    for (int m = 0; m < 3; m++) s += Ta(m, r) * (Tb(m, 3) - Ta(m, 3));
    R(r, 3) = s;
  }
  return R;
}

// lower Cholesky of a symmetric 6x6, transposed: LLT(cov^-1).matrixL().transpose() (optimization_be.cpp:922-923)
inline bool sqrt_info_from_cov(const double* cov /*36 row-major*/, double* out /*36*/) {
  // invert via Cholesky of cov, then Cholesky of the inverse
  double L[36] = {0}, X[36] = {0}, P[36];
  for (int c = 0; c < 6; c++)
    for (int r = c; r < 6; r++) {
      double s = cov[6 * r + c];
      for (int m = 0; m < c; m++) s -= L[6 * r + m] * L[6 * c + m];
      if (r == c) {
        if (!(s > 0)) return false;
        L[6 * c + c] = std::sqrt(s);
      } else {
        L[6 * r + c] = s / L[6 * c + c];
      }
    }
  for (int c = 0; c < 6; c++)
    for (int r = c; r < 6; r++) {
      double s = (r == c) ? 1.0 : 0.0;
      for (int m = c; m < r; m++) s -= L[6 * r + m] * X[6 * m + c];
      X[6 * r + c] = s / L[6 * r + r];
    }
  for (int r = 0; r < 6; r++)
    for (int c = 0; c < 6; c++) {
      double s = 0;
      for (int m = std::max(r, c); m < 6; m++) s += X[6 * m + r] * X[6 * m + c];
      P[6 * r + c] = s;
    }
  double M[36] = {0};
  for (int c = 0; c < 6; c++)
    for (int r = c; r < 6; r++) {
      double s = P[6 * r + c];
      for (int m = 0; m < c; m++) s -= M[6 * r + m] * M[6 * c + m];
      if (r == c) {
        if (!(s > 0)) return false;
        M[6 * c + c] = std::sqrt(s);
      } else {
        M[6 * r + c] = s / M[6 * c + c];
      }
    }
  for (int r = 0; r < 6; r++)
    for (int c = 0; c < 6; c++) out[6 * r + c] = M[6 * c + r];
  return true;
}

// Flattened problem with owning storage + the cvb_ba_problem view
struct Flat {
  std::vector<double> pose, sb, extr, intr, dist, lm, obs_sigma, imu_dt, imu_acc, imu_gyr, imu_acc0, imu_gyr0, edge_q, edge_t,
      edge_S, cam_xi;
  std::vector<int32_t> cam_model, dist_model;
  std::vector<float> obs_uv;
  std::vector<uint8_t> pose_const, edge_robust;
  std::vector<int32_t> cam_of_kf, lm_obs_ptr, obs_kf, imu_i, imu_j, imu_ptr, edge_i, edge_j;
  double imu_noise[5] = {0, 0, 0, 0, 9.81};
  cvb_ba_problem view() const {
    cvb_ba_problem p{};
    p.K = (int32_t)pose_const.size();
    p.L = (int32_t)(lm.size() / 3);
    p.n_obs = (int32_t)obs_kf.size();
    p.n_imu = (int32_t)imu_i.size();
    p.n_edge = (int32_t)edge_i.size();
    p.n_cam = (int32_t)(extr.size() / 7);
    p.pose = pose.data(); p.speedbias = sb.data(); p.pose_const = pose_const.data(); p.cam_of_kf = cam_of_kf.data();
    p.extr = extr.data(); p.intr = intr.data(); p.dist = dist.data(); p.lm = lm.data(); p.lm_obs_ptr = lm_obs_ptr.data();
    p.obs_kf = obs_kf.data(); p.obs_uv = obs_uv.data(); p.obs_sigma = obs_sigma.data(); p.obs_skip = nullptr;
    p.imu_i = imu_i.data(); p.imu_j = imu_j.data(); p.imu_ptr = imu_ptr.data(); p.imu_dt = imu_dt.data();
    p.imu_acc = imu_acc.data(); p.imu_gyr = imu_gyr.data(); p.imu_acc0 = imu_acc0.data(); p.imu_gyr0 = imu_gyr0.data();
    p.imu_noise = imu_noise;
    p.edge_i = edge_i.data(); p.edge_j = edge_j.data(); p.edge_q = edge_q.data(); p.edge_t = edge_t.data();
    p.edge_sqrt_info = edge_S.data(); p.edge_robust = edge_robust.data();
    if (cam_model.size() == (size_t)p.n_cam) { p.cam_model = cam_model.data(); p.dist_model = dist_model.data(); p.cam_xi = cam_xi.data(); }
    return p;
  }
};

}  // namespace detail

// parameters the reference reads from covins_params (config/config_backend.yaml)
struct OptParams {
  bool gba_fix_poses_loaded_maps = false;       // opt.gba_fix_poses_loaded_maps (optimization_be.cpp:338)
  bool gba_use_map_loop_constraints = true;     // :539
  double th_gba_outlier_global = 0.92;          // :277
  bool pgo_fix_kfs_after_gba = true;            // :875
  bool pgo_fix_poses_loaded_maps = true;        // :878
  int pgo_iteration_limit = 10;                 // :1029
  bool use_nbr_kfs = true;                      // :976
  bool use_robust_loss = true;                  // :934
  double robust_loss_th = 0.5;                  // :840
  double wt_kf_r = 10.0, wt_kf_t = 1.0, wt_kf_n1 = 10.0, wt_kf_n23 = 2.0, wt_kf_n45 = 3.0;   // :896-903
  bool placerec_type_covins = false;            // placerec.type == "COVINS" (:929)
  float img_match_thres = 40.0f, ratio_thres = 0.8f;    // features (placerec_gen_be.cpp:107-108)
  int matches_thres = 25, matches_thres_merge = 25;     // placerec (placerec_gen_be.cpp:118-121)
};

// ---------------------------------------------------------------------------------------------------------------
// Optimization::GlobalBundleAdjustment — same signature meaning as optimization_be.hpp:38-40.
// ---------------------------------------------------------------------------------------------------------------
template <class MapPtr>
void GlobalBundleAdjustment(Context& ctx, MapPtr map, int interations_limit, double /*time_limit*/, bool visual_only = false,
                            bool outlier_removal = true, bool /*estimate_bias*/ = false, const OptParams& P = OptParams()) {
  using KeyframePtr = typename std::decay<decltype(map->GetKeyframesVec()[0])>::type;
  using KF = typename KeyframePtr::element_type;
  using Transform = typename std::decay<decltype(map->GetKeyframesVec()[0]->GetPoseTws())>::type;
  std::printf("+++ GBA: Start +++\n");
  auto keyframes = map->GetKeyframesVec();   // id-sorted std::map order (map_base.cpp:63-69)
  auto landmarks = map->GetLandmarksVec();
  std::printf("--> KFs: %zu\n--> LMs: %zu\n", keyframes.size(), landmarks.size());

  detail::Flat F;
  std::map<const KF*, int> kf_index;
  std::vector<KeyframePtr> kfs;               // valid keyframes, canonical order
  for (auto& kf : keyframes) {
    if (kf->IsInvalid()) continue;
    kf_index[kf.get()] = (int)kfs.size();
    kfs.push_back(kf);
  }
  const int K = (int)kfs.size();
  F.pose.resize(7 * (size_t)K); F.sb.resize(9 * (size_t)K); F.extr.resize(7 * (size_t)K); F.intr.resize(4 * (size_t)K);
  F.dist.resize(4 * (size_t)K); F.pose_const.assign(K, 0); F.cam_of_kf.resize(K);
  F.imu_ptr.push_back(0);
  for (int k = 0; k < K; k++) {
    auto& kf = kfs[k];
    // UpdateCeresFromState (keyframe_base.cpp:486-521) restated on the flat arrays
    kf->UpdateCeresFromState(&F.pose[7 * (size_t)k], &F.sb[9 * (size_t)k], &F.extr[7 * (size_t)k]);
    F.cam_of_kf[k] = k;
    if (kf->id_.first == 0 && kf->id_.second == map->id_map_) F.pose_const[k] = 1;                 // :88-89, 329-331
    if (kf->is_loaded_ && P.gba_fix_poses_loaded_maps) F.pose_const[k] = 1;                        // :338-341
    if (!Adapter<KF>::camera(*kf, &F.intr[4 * (size_t)k], &F.dist[4 * (size_t)k])) {
      std::printf("FATAL: Unknown projection type.\n");                                            // :112-114
      std::exit(-1);
    }
    {
      int cm = 0, dm = 0; double xi = 0.0;
      Adapter<KF>::camera_model(*kf, &cm, &dm, &xi);                                               // :186-231
      F.cam_model.resize(F.pose_const.size(), 0); F.dist_model.resize(F.pose_const.size(), 0); F.cam_xi.resize(F.pose_const.size(), 0.0);
      F.cam_model[k] = cm; F.dist_model[k] = dm; F.cam_xi[k] = xi;
    }
    if (!visual_only) {                                                                             // :117-144, 367-421
      auto pred = kf->GetPredecessor();
      if (!pred || pred->IsInvalid()) {
        if (kf->id_.first != 0) {
          std::printf("FATAL: KF %zu|%zu: no predecessor\n", (size_t)kf->id_.first, (size_t)kf->id_.second);
          std::exit(-1);
        }
        continue;
      }
      if (Adapter<KF>::imu_count(*kf) == 0) {
        std::printf("KF %zu|%zu 0 IMU measurements - skip IMU factor\n", (size_t)kf->id_.first, (size_t)kf->id_.second);   // :382-385
        continue;
      }
      std::vector<double> dt, acc, gyr;
      double a0[3], g0[3];
      Adapter<KF>::imu_samples(*kf, dt, acc, gyr, a0, g0, F.imu_noise);
      F.imu_i.push_back(kf_index.at(pred.get()));
      F.imu_j.push_back(k);
      F.imu_dt.insert(F.imu_dt.end(), dt.begin(), dt.end());
      F.imu_acc.insert(F.imu_acc.end(), acc.begin(), acc.end());
      F.imu_gyr.insert(F.imu_gyr.end(), gyr.begin(), gyr.end());
      F.imu_acc0.insert(F.imu_acc0.end(), a0, a0 + 3);
      F.imu_gyr0.insert(F.imu_gyr0.end(), g0, g0 + 3);
      F.imu_ptr.push_back((int32_t)F.imu_dt.size());
    }
  }
  // landmarks + observations (canonical: observations sorted by keyframe index; the reference iterates a
  // pointer-ordered std::map, typedefs_base.hpp:187)
  struct ObsRef { KeyframePtr kf; int lm; int feat; };
  std::vector<ObsRef> obs_ref;
  using LandmarkPtr = typename std::decay<decltype(landmarks[0])>::type;
  std::vector<LandmarkPtr> lms;
  F.lm_obs_ptr.push_back(0);
  for (auto& lm : landmarks) {
    if (lm->IsInvalid()) continue;
    const auto observations = lm->GetObservations();
    std::vector<std::pair<int, int>> ob;   // (kf index, feature id)
    for (const auto& mit : observations) {
      auto kfx = mit.first;
      if (!kfx || kfx->IsInvalid()) continue;
      ob.emplace_back(kf_index.at(kfx.get()), (int)mit.second);
    }
    std::sort(ob.begin(), ob.end());
    const auto pos = lm->GetWorldPos();
    F.lm.push_back(pos[0]); F.lm.push_back(pos[1]); F.lm.push_back(pos[2]);
    for (auto& o : ob) {
      auto& kfx = kfs[o.first];
      const auto& kp = kfx->keypoints_distorted_[o.second];
      F.obs_kf.push_back(o.first);
      F.obs_uv.push_back((float)kp[0]); F.obs_uv.push_back((float)kp[1]);
      F.obs_sigma.push_back((kfx->keypoints_aors_[o.second][1] + 1) * 2.0);                         // :183-184, 477-478
      obs_ref.push_back({kfx, (int)lms.size(), o.second});
    }
    F.lm_obs_ptr.push_back((int32_t)F.obs_kf.size());
    lms.push_back(lm);
  }
  // loop edges (:236-254, 532-557): sqrt_info = diag(100 I3, 1e4 I3)
  if (P.gba_use_map_loop_constraints) {
    for (const auto& lc : map->GetLoopConstraints()) {
      auto i1 = kf_index.find(lc.kf1.get()), i2 = kf_index.find(lc.kf2.get());
      if (i1 == kf_index.end() || i2 == kf_index.end()) {
        std::printf("WARN: Loop KF missing -- skip loop\n");                                        // :546-549
        continue;
      }
      double p7[7];
      detail::transform_to_pose7(lc.T_s1_s2, p7);
      F.edge_i.push_back(i1->second); F.edge_j.push_back(i2->second);
      F.edge_q.insert(F.edge_q.end(), p7, p7 + 4);
      F.edge_t.insert(F.edge_t.end(), p7 + 4, p7 + 7);
      for (int r = 0; r < 6; r++)
        for (int c = 0; c < 6; c++) F.edge_S.push_back(r == c ? (r < 3 ? 100.0 : 1e4) : 0.0);
      F.edge_robust.push_back(1);
    }
  }
  const cvb_ba_problem prob = F.view();
  std::vector<double> o_pose(7 * (size_t)K), o_sb(9 * (size_t)K), o_lm(F.lm.size());
  std::vector<int32_t> owner(lms.size() ? lms.size() : 1);
  std::vector<uint8_t> removed(F.obs_kf.size() ? F.obs_kf.size() : 1, 0);
  cvb_ba_result res{};
  res.pose = o_pose.data(); res.speedbias = o_sb.data(); res.lm = o_lm.data(); res.lm_owner = owner.data();
  cvb_gba_options g{interations_limit, visual_only ? 1 : 0, outlier_removal ? 1 : 0, P.th_gba_outlier_global};
  ctx.check(cvb_gba(ctx.get(), &prob, &g, &res, removed.data()), "cvb_gba");

  // round-1 outlier purge, written into the map exactly like optimization_be.cpp:282-288
  size_t num_bad = 0;
  for (size_t i = 0; i < obs_ref.size(); i++)
    if (removed[i]) {
      obs_ref[i].kf->EraseLandmark(obs_ref[i].feat);
      lms[obs_ref[i].lm]->EraseObservation(obs_ref[i].kf);
      ++num_bad;
    }
  if (outlier_removal) std::printf("--> GBA removed %zu of %zu observations\n", num_bad, 2 * obs_ref.size());
  // Recover optimized data — keyframes (:572-595)
  for (int k = 0; k < K; k++) {
    auto& kf = kfs[k];
    kf->SetPoseTws(detail::pose7_to_transform<Transform>(&o_pose[7 * (size_t)k]));
    kf->SetPoseOptimized();
    if (!visual_only) {
      const double* s = &o_sb[9 * (size_t)k];
      kf->SetStateBias({s[3], s[4], s[5]}, {s[6], s[7], s[8]});
      kf->SetStateVelocity({s[0], s[1], s[2]});
      kf->SetVelBiasOptimized();
    }
    kf->is_gba_optimized_ = true;
  }
  // landmarks (:598-609): only those that were in the problem
  for (size_t l = 0; l < lms.size(); l++) {
    if (owner[l] < 0) continue;
    lms[l]->SetWorldPos({o_lm[3 * l], o_lm[3 * l + 1], o_lm[3 * l + 2]});
    lms[l]->SetOptimized();
    lms[l]->is_gba_optimized_ = true;
  }
  std::printf("--> Clean Map\n");
  map->Clean();                                                                                     // :614
  std::printf("--> done.\n+++ GBA: End +++\n");
}

// ---------------------------------------------------------------------------------------------------------------
// Optimization::PoseGraphOptimization — optimization_be.hpp:46-47.  PoseMap = std::map<idpair, Transform>.
// ---------------------------------------------------------------------------------------------------------------
template <class MapPtr, class PoseMap>
void PoseGraphOptimization(Context& ctx, MapPtr map, PoseMap corrected_poses, const OptParams& P = OptParams()) {
  using KeyframePtr = typename std::decay<decltype(map->GetKeyframesVec()[0])>::type;
  using KF = typename KeyframePtr::element_type;
  using Transform = typename std::decay<decltype(map->GetKeyframesVec()[0]->GetPoseTws())>::type;
  auto keyframes = map->GetKeyframesVec();
  auto landmarks = map->GetLandmarksVec();
  detail::Flat F;
  std::map<const KF*, int> kf_index;
  std::vector<KeyframePtr> kfs;
  for (auto& kf : keyframes) {
    if (kf->IsInvalid()) continue;
    kf_index[kf.get()] = (int)kfs.size();
    kfs.push_back(kf);
  }
  const int K = (int)kfs.size();
  F.pose.resize(7 * (size_t)K); F.sb.assign(9 * (size_t)K, 0.0); F.extr.resize(7 * (size_t)K);
  F.intr.assign(4 * (size_t)K, 1.0); F.dist.assign(4 * (size_t)K, 0.0); F.pose_const.assign(K, 0); F.cam_of_kf.resize(K);
  F.lm_obs_ptr.push_back(0); F.imu_ptr.push_back(0);
  for (int k = 0; k < K; k++) {
    auto& kf = kfs[k];
    double tmp_pose[7];
    kf->UpdateCeresFromState(tmp_pose, &F.sb[9 * (size_t)k], &F.extr[7 * (size_t)k]);
    auto mit = corrected_poses.find(kf->id_);                                                        // :854-868
    const Transform T_ws_init = (mit != corrected_poses.end()) ? mit->second : kf->GetPoseTws();
    detail::transform_to_pose7(T_ws_init, &F.pose[7 * (size_t)k]);
    F.cam_of_kf[k] = k;
    if (kf->id_.first == 0 && kf->id_.second == map->id_map_) F.pose_const[k] = 1;                  // :870-871
    if (kf->is_gba_optimized_ && P.pgo_fix_kfs_after_gba) F.pose_const[k] = 1;                      // :875-877
    else if (kf->is_loaded_ && P.pgo_fix_poses_loaded_maps) F.pose_const[k] = 1;                    // :878-881
  }
  double S1[36] = {0}, S23[36] = {0}, S45[36] = {0};
  for (int d = 0; d < 6; d++) {
    S1[7 * d] = (d < 3 ? P.wt_kf_r : P.wt_kf_t) * P.wt_kf_n1;                                        // :896-898
    S23[7 * d] = S1[7 * d] / P.wt_kf_n23;
    S45[7 * d] = S1[7 * d] / P.wt_kf_n45;
  }
  auto push_edge = [&](int i, int j, const Transform& T12, const double* S, bool robust) {
    double p7[7];
    detail::transform_to_pose7(T12, p7);
    F.edge_i.push_back(i); F.edge_j.push_back(j);
    F.edge_q.insert(F.edge_q.end(), p7, p7 + 4);
    F.edge_t.insert(F.edge_t.end(), p7 + 4, p7 + 7);
    F.edge_S.insert(F.edge_S.end(), S, S + 36);
    F.edge_robust.push_back(robust ? 1 : 0);
  };
  for (const auto& lc : map->GetLoopConstraints()) {                                                 // :910-943
    double Sl[36];
    if (P.placerec_type_covins) {
      std::copy(S1, S1 + 36, Sl);
    } else {
      double cov[36];
      for (int r = 0; r < 6; r++)
        for (int c = 0; c < 6; c++) cov[6 * r + c] = lc.cov_mat(r, c);
      if (!detail::sqrt_info_from_cov(cov, Sl)) {
        std::printf("FATAL: loop covariance is not positive definite\n");
        std::exit(-1);
      }
    }
    push_edge(kf_index.at(lc.kf1.get()), kf_index.at(lc.kf2.get()), lc.T_s1_s2, Sl, P.use_robust_loss);
  }
  std::set<std::pair<const KF*, const KF*>> inserted_edges;                                          // :907
  for (int k = 0; k < K; k++) {                                                                      // successor edges :947-972
    auto& kf = kfs[k];
    auto succ = kf->GetSuccessor();
    if (!succ) continue;
    auto key = std::make_pair((const KF*)kf.get(), (const KF*)succ.get());
    if (inserted_edges.count(key)) {
      std::printf("WARN: KF edge already added\n");
      continue;
    }
    inserted_edges.insert(key);
    push_edge(k, kf_index.at(succ.get()), detail::rel_transform(kf->GetPoseTws_vio(), succ->GetPoseTws_vio()), S1, false);
  }
  if (P.use_nbr_kfs) {                                                                               // :976-1021
    for (int k = 0; k < K; k++) {
      auto& kf = kfs[k];
      std::vector<KeyframePtr> connections;
      KeyframePtr temp_kf = kf;
      for (int j = 1; j < 6; ++j)
        if (int(kf->id_.first) - j > 0) {
          temp_kf = temp_kf->GetPredecessor();
          connections.push_back(temp_kf);
        }
      size_t n = 0;
      for (auto& kfc : connections) {
        n++;
        const double* S = (n <= 1) ? S1 : (n <= 3 ? S23 : S45);
        auto key = std::make_pair((const KF*)kf.get(), (const KF*)kfc.get());
        if (inserted_edges.count(key)) continue;
        inserted_edges.insert(key);
        push_edge(k, kf_index.at(kfc.get()), detail::rel_transform(kf->GetPoseTws_vio(), kfc->GetPoseTws_vio()), S, false);
      }
    }
  }
  const cvb_ba_problem prob = F.view();
  cvb_ba_options o{};
  o.max_iterations = P.pgo_iteration_limit;
  o.visual_only = 1;
  o.cauchy_reproj = 0.0;
  o.cauchy_edge = P.robust_loss_th;
  o.world = 1;
  std::vector<double> o_pose(7 * (size_t)K);
  cvb_ba_result res{};
  res.pose = o_pose.data();
  ctx.check(cvb_ba_solve(ctx.get(), &prob, &o, &res), "cvb_ba_solve(PGO)");

  // Recover the optimized data (:1033-1051)
  std::map<typename std::decay<decltype(kfs[0]->id_)>::type, Transform> non_corrected_poses;
  for (int k = 0; k < K; k++) {
    auto& kf = kfs[k];
    const Transform T_ws_uncorrected = kf->GetPoseTws();
    non_corrected_poses[kf->id_] = T_ws_uncorrected;
    const Transform T_ws_corrected = detail::pose7_to_transform<Transform>(&o_pose[7 * (size_t)k]);
    const auto vel = kf->GetStateVelocity();
    kf->SetPoseTws(T_ws_corrected);
    double v[3];   // R_corr * R_uncorr^T * v  (:1046-1047)
    for (int r = 0; r < 3; r++) {
      double s = 0;
      for (int c = 0; c < 3; c++) {
        double m = 0;
        for (int x = 0; x < 3; x++) m += T_ws_corrected(r, x) * T_ws_uncorrected(c, x);
        s += m * vel[c];
      }
      v[r] = s;
    }
    kf->SetStateVelocity({v[0], v[1], v[2]});
    kf->SetPoseOptimized();
  }
  // Landmarks re-anchored through their reference keyframe (:1054-1083)
  for (auto& lm : landmarks) {
    if (lm->IsInvalid()) continue;
    auto kf_ref = lm->GetReferenceKeyframe();
    if (!kf_ref) {
      if (!lm->GetObservations().empty()) map->EraseLandmark(lm);
      continue;
    }
    auto mit = non_corrected_poses.find(kf_ref->id_);
    if (mit == non_corrected_poses.end()) {
      map->EraseLandmark(lm);
      continue;
    }
    const Transform& Tu = mit->second;
    const auto pw = lm->GetWorldPos();
    double ps[3], pc[3];
    for (int r = 0; r < 3; r++) {
      double s = 0;
      for (int c = 0; c < 3; c++) s += Tu(c, r) * (pw[c] - Tu(c, 3));
      ps[r] = s;
    }
    const Transform Tc = kf_ref->GetPoseTws();
    for (int r = 0; r < 3; r++) pc[r] = Tc(r, 0) * ps[0] + Tc(r, 1) * ps[1] + Tc(r, 2) * ps[2] + Tc(r, 3);
    lm->SetWorldPos({pc[0], pc[1], pc[2]});
    lm->SetOptimized();
  }
  std::printf("--> PGO END \n");
}

// ---------------------------------------------------------------------------------------------------------------
// Matching blocks
// ---------------------------------------------------------------------------------------------------------------
struct Match {   // covins::Match (include/covins/matcher/MatchingAlgorithm.h:56-70)
  size_t idxA, idxB;
  float distance;
};
using Matches = std::vector<Match>;

// The ORB branch of the candidate loop of PlaceRecognitionG::ComputeSE3 (placerec_gen_be.cpp:72-125) for ALL
// candidates at once: knnMatch(query, cand, 2) + distance/ratio filter.  descriptors are row-major [n][32] uint8
// (cv::Mat descriptors_add_ rows).  Returns img_matches per candidate; discarded[i] is set as at :118-124.
inline std::vector<Matches> MatchCandidatesORB(Context& ctx, const uint8_t* query, int n_query,
                                               const std::vector<const uint8_t*>& cand_desc, const std::vector<int>& cand_rows,
                                               const std::vector<bool>& same_client, const OptParams& P,
                                               std::vector<bool>* discarded) {
  const int n_seg = (int)cand_desc.size();
  std::vector<int32_t> seg(n_seg + 1, 0);
  for (int s = 0; s < n_seg; s++) seg[s + 1] = seg[s] + cand_rows[s];
  std::vector<uint8_t> train((size_t)seg[n_seg] * 32);
  for (int s = 0; s < n_seg; s++) std::copy(cand_desc[s], cand_desc[s] + (size_t)cand_rows[s] * 32, train.begin() + (size_t)seg[s] * 32);
  std::vector<int32_t> mt((size_t)n_seg * n_query), nm(n_seg);
  std::vector<float> md((size_t)n_seg * n_query);
  ctx.check(cvb_match_hamming_batch(ctx.get(), query, n_query, train.data(), seg.data(), n_seg, P.img_match_thres, P.ratio_thres,
                                    mt.data(), md.data(), nm.data()),
            "cvb_match_hamming_batch");
  std::vector<Matches> out(n_seg);
  if (discarded) discarded->assign(n_seg, false);
  for (int s = 0; s < n_seg; s++) {
    for (int q = 0; q < n_query; q++) {
      const int32_t t = mt[(size_t)s * n_query + q];
      if (t >= 0) out[s].push_back(Match{(size_t)q, (size_t)t, md[(size_t)s * n_query + q]});
    }
    const int nmatches = (int)out[s].size();
    if (discarded) {
      if (same_client[s] && nmatches < P.matches_thres) (*discarded)[s] = true;        // placerec_gen_be.cpp:118-120
      else if (nmatches < P.matches_thres_merge) (*discarded)[s] = true;               // :121-123
    }
  }
  return out;
}

// The DenseMatcher block of PlaceRecognition::ComputeSE3 (placerec_be.cpp:85-91) for all candidates at once.
inline std::vector<Matches> LandmarkMatchCandidates(Context& ctx, const uint8_t* query, const uint8_t* skip_query, int n_query,
                                                    const std::vector<const uint8_t*>& cand_desc,
                                                    const std::vector<const uint8_t*>& cand_skip, const std::vector<int>& cand_rows,
                                                    float distance_threshold = 50.0f, int num_best = 4) {
  const int n_seg = (int)cand_desc.size();
  std::vector<int32_t> seg(n_seg + 1, 0);
  for (int s = 0; s < n_seg; s++) seg[s + 1] = seg[s] + cand_rows[s];
  const size_t rows = (size_t)seg[n_seg];
  std::vector<uint8_t> B(rows * 32), skipB(rows);
  for (int s = 0; s < n_seg; s++) {
    std::copy(cand_desc[s], cand_desc[s] + (size_t)cand_rows[s] * 32, B.begin() + (size_t)seg[s] * 32);
    std::copy(cand_skip[s], cand_skip[s] + cand_rows[s], skipB.begin() + seg[s]);
  }
  std::vector<int32_t> oA(rows ? rows : 1), oB(rows ? rows : 1), n(n_seg);
  std::vector<float> oD(rows ? rows : 1);
  ctx.check(cvb_landmark_match_batch(ctx.get(), query, skip_query, n_query, B.data(), skipB.data(), seg.data(), n_seg,
                                     distance_threshold, num_best, oA.data(), oB.data(), oD.data(), n.data()),
            "cvb_landmark_match_batch");
  std::vector<Matches> out(n_seg);
  for (int s = 0; s < n_seg; s++)
    for (int m = 0; m < n[s]; m++) out[s].push_back(Match{(size_t)oA[seg[s] + m], (size_t)oB[seg[s] + m], oD[seg[s] + m]});
  return out;
}

// Resident-map descriptor database (cvb_db_*): the ORB descriptors of the map's keyframes live in HBM; the candidate
// loop of PlaceRecognitionG::ComputeSE3 (placerec_gen_be.cpp:60-135) becomes one call per query keyframe.  The database
// index of a keyframe is its insertion order; keep it next to the keyframe (e.g. std::map<idpair, int>).
class DescriptorDatabase {
 public:
  explicit DescriptorDatabase(Context& ctx) : ctx_(ctx) { ctx_.check(cvb_db_create(ctx_.get(), 32, &db_), "cvb_db_create"); }
  ~DescriptorDatabase() { if (db_) cvb_db_destroy(ctx_.get(), db_); }
  DescriptorDatabase(const DescriptorDatabase&) = delete;
  DescriptorDatabase& operator=(const DescriptorDatabase&) = delete;
  // descriptors: kf->descriptors_add_ (CV_8U, continuous, rows x 32; keyframe_be.cpp:103,137) → returns the database index
  int AddKeyframe(const uint8_t* descriptors, int rows) {
    const int32_t r = rows;
    ctx_.check(cvb_db_append(ctx_.get(), db_, descriptors, &r, 1), "cvb_db_append");
    return n_kf_++;
  }
  int size() const { return n_kf_; }
  // knnMatch(k=2) + distance/ratio filter of the query keyframe against EVERY keyframe of the database:
  // result[db index] == the reference's img_matches for that candidate (accepted queries ascending, :102-114)
  std::vector<Matches> MatchAll(const uint8_t* query, int n_query, const OptParams& P) {
    std::vector<int32_t> nm(n_kf_ > 0 ? n_kf_ : 1);
    if (cap_ == 0) cap_ = 4096;
    for (;;) {
      m_kf_.resize(cap_); m_q_.resize(cap_); m_t_.resize(cap_); m_d_.resize(cap_);
      int32_t total = 0;
      ctx_.check(cvb_db_match_hamming(ctx_.get(), db_, query, n_query, P.img_match_thres, P.ratio_thres, nm.data(), m_kf_.data(),
                                      m_q_.data(), m_t_.data(), m_d_.data(), cap_, &total),
                 "cvb_db_match_hamming");
      if (total <= cap_) {
        std::vector<Matches> out(n_kf_);
        for (int i = 0; i < total; i++) out[m_kf_[i]].push_back(Match{(size_t)m_q_[i], (size_t)m_t_[i], m_d_[i]});
        return out;
      }
      cap_ = total + total / 4 + 16;
    }
  }

 private:
  Context& ctx_;
  cvb_db* db_ = nullptr;
  int n_kf_ = 0, cap_ = 0;
  std::vector<int32_t> m_kf_, m_q_, m_t_;
  std::vector<float> m_d_;
};

// Landmark::ComputeDescriptor (landmark_be.cpp:49-92) for a batch of landmarks: cand[l] = the descriptor rows
// (kf->descriptors_.row(feat_idx), 32 bytes each) of the landmark's valid observers in observation order.
// Returns per landmark the index of the chosen observer (-1: no observer, descriptor unchanged) and writes the chosen
// descriptor to out_desc[l] (32 bytes each; pass the current descriptors in, as the reference keeps them on early return).
inline std::vector<int> ComputeLandmarkDescriptors(Context& ctx, const std::vector<std::vector<const uint8_t*>>& cand,
                                                   uint8_t* out_desc) {
  const int n_lm = (int)cand.size();
  std::vector<int32_t> ptr(n_lm + 1, 0);
  for (int l = 0; l < n_lm; l++) ptr[l + 1] = ptr[l] + (int32_t)cand[l].size();
  std::vector<uint8_t> rows((size_t)ptr[n_lm] * 32 + 32);
  for (int l = 0; l < n_lm; l++)
    for (size_t j = 0; j < cand[l].size(); j++) std::copy(cand[l][j], cand[l][j] + 32, rows.begin() + ((size_t)ptr[l] + j) * 32);
  std::vector<int32_t> best(n_lm > 0 ? n_lm : 1);
  ctx.check(cvb_landmark_descriptor_batch(ctx.get(), rows.data(), ptr.data(), n_lm, best.data(), out_desc), "cvb_landmark_descriptor_batch");
  return std::vector<int>(best.begin(), best.begin() + n_lm);
}

}  // namespace covins_b200

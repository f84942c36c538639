// oracle/ref/cereal_fixture_gen.cpp — TEST INFRASTRUCTURE ONLY (fixture writer; never linked into covins_b200/).
//
// Writes a small COVINS map directory with the REAL serialisation stack of the reference:
//   * the vendored cereal (covins_comm/thirdparty/cereal) BinaryOutputArchive,
//   * the reference's own message types and `save` templates, included from where they lie:
//       covins_comm/include/covins/covins_base/msgs/msg_keyframe.hpp   (MsgKeyframe file-save branch :129-143,
//                                                                       Eigen save :211-221, cv::Mat save :237-262,
//                                                                       PreintegrationData :37-42)
//       covins_comm/include/covins/covins_base/msgs/msg_landmark.hpp   (MsgLandmark file-save branch :69-73)
//       covins_comm/include/covins/covins_base/typedefs_base.hpp       (VICalibration::serialize :376-380)
//     with stand-in Eigen::Matrix / cv::Mat types (oracle/ref/shim/) in place of the absent Eigen / OpenCV headers,
//   * the write sequence of Map::SaveToFile (covins_backend/src/covins_backend/map_be.cpp:861-908): one archive per
//     keyframe / landmark / map-data file, `oarchive(msg)` into a stringstream, the string written to the file.
// MsgMap lives in map_be.hpp (which needs the whole backend); its 5-field serialize (map_be.hpp:126-136) is restated
// below — the byte layout of its members is still produced by cereal + the reference's Eigen save template.
//
// Output: <out>/keyframes/keyframes<i>.txt, <out>/mappoints/mappoints<i>.txt, <out>/mapdata.txt and <out>/manifest.json
// (the field values that went in, for tests/test_mapio_cereal.py).  Usage: cereal_fixture_gen <out_dir>
#include <sys/stat.h>

#include <cstdio>
#include <fstream>
#include <iomanip>
#include <sstream>
#include <string>

#include "covins_base/msgs/msg_keyframe.hpp"
#include "covins_base/msgs/msg_landmark.hpp"

// (the out-of-line members of the two message types — constructors, SetMsgType — are compiled where they lie by the
//  Makefile: covins_comm/src/covins_base/msgs/msg_keyframe.cpp, msg_landmark.cpp)

namespace {

using covins::TypeDefs::idpair;

struct MsgMapRestated {   // covins_backend/include/covins/covins_backend/map_be.hpp:126-136
  size_t id_map;
  std::vector<covins::TypeDefs::idpair> keyframes1;
  std::vector<covins::TypeDefs::idpair> keyframes2;
  std::vector<covins::TypeDefs::TransformType> transforms12;
  std::vector<covins::TypeDefs::Matrix6Type> cov;
  template <class Archive>
  void serialize(Archive& archive) { archive(id_map, keyframes1, keyframes2, transforms12, cov); }
};

// deterministic pseudo-values (exactly representable sums of small dyadic fractions → the JSON round-trips bit-exactly)
double val(int a, int b, int c = 0) { return 0.125 * a - 0.5 * b + 0.03125 * c + 1.0; }

template <class M>
void fill(M& m, int seed) {
  for (int i = 0; i < m.rows(); i++)
    for (int j = 0; j < m.cols(); j++) m(i, j) = (typename std::remove_reference<decltype(m(0, 0))>::type)val(seed, i, j * 3 + 1);
}

struct Json {
  std::ostringstream s;
  bool first = true;
  Json() { s << std::setprecision(17); }
  void key(const std::string& k) { s << (first ? "" : ",") << "\"" << k << "\":"; first = false; }
  template <class M>
  void mat(const std::string& k, const M& m) {   // row-major nested list
    key(k);
    s << "[";
    for (int i = 0; i < m.rows(); i++) {
      s << (i ? "," : "") << "[";
      for (int j = 0; j < m.cols(); j++) s << (j ? "," : "") << (double)m(i, j);
      s << "]";
    }
    s << "]";
  }
  void num(const std::string& k, double v) { key(k); s << v; }
  void pair(const std::string& k, const idpair& p) { key(k); s << "[" << p.first << "," << p.second << "]"; }
  void vec(const std::string& k, const std::vector<double>& v) {
    key(k);
    s << "[";
    for (size_t i = 0; i < v.size(); i++) s << (i ? "," : "") << v[i];
    s << "]";
  }
  template <class V>
  void matvec(const std::string& k, const V& v) {   // vector of fixed-size Eigen → list of row-major flattened lists
    key(k);
    s << "[";
    for (size_t e = 0; e < v.size(); e++) {
      s << (e ? "," : "") << "[";
      for (int i = 0; i < v[e].rows(); i++)
        for (int j = 0; j < v[e].cols(); j++) s << ((i || j) ? "," : "") << (double)v[e](i, j);
      s << "]";
    }
    s << "]";
  }
  void cvmat(const std::string& k, const cv::Mat& m) {
    key(k);
    s << "{\"rows\":" << m.rows << ",\"cols\":" << m.cols << ",\"type\":" << m.type() << ",\"data\":[";
    const size_t n = (size_t)m.rows * m.cols;
    for (size_t i = 0; i < n; i++) {
      if (i) s << ",";
      if ((m.type() & 7) == 5) s << (double)reinterpret_cast<const float*>(m.ptr())[i];
      else s << (int)m.ptr()[i];
    }
    s << "]}";
  }
};

template <class Msg>
void write_archive(const std::string& path, const Msg& msg) {   // map_be.cpp:866-876
  std::ofstream fs;
  fs.open(path);
  std::stringstream ss;
  {
    cereal::BinaryOutputArchive oarchive(ss);
    oarchive(msg);
  }
  fs << ss.str();
  fs.close();
}

}  // namespace

int main(int argc, char** argv) {
  if (argc < 2) { std::fprintf(stderr, "usage: %s <out_dir>\n", argv[0]); return 2; }
  const std::string out = argv[1];
  mkdir(out.c_str(), 0777);
  mkdir((out + "/keyframes").c_str(), 0777);
  mkdir((out + "/mappoints").c_str(), 0777);
  std::ostringstream manifest;
  manifest << "{\"keyframes\":[";
  const idpair kf_ids[3] = {{0, 0}, {1, 0}, {0, 1}};
  for (int k = 0; k < 3; k++) {
    covins::MsgKeyframe msg;
    msg.save_to_file = true;                       // keyframe_be.cpp:304
    msg.id = kf_ids[k];
    msg.timestamp = 100.0 + 0.25 * k;
    Eigen::Matrix4d Tsc; fill(Tsc, 10 + k);
    Eigen::VectorXd dc(4, 1); fill(dc, 20 + k);
    Eigen::Vector3d a0; fill(a0, 30 + k);
    msg.calibration = covins::VICalibration(Tsc, covins::eCamModel::PINHOLE, k == 2 ? covins::eDistortionModel::EQUI : covins::eDistortionModel::RADTAN,
                                            dc, 752.0, 480.0, 458.654, 457.296, 367.215, 248.375, 176.0, 7.8, 2e-3, 1.7e-4, 0.03,
                                            0.0083, 3e-3, 2e-5, 3600.0, 9.81, a0, 200, 0.0, 0.5);
    msg.img_dim_x_min = -3 + k; msg.img_dim_y_min = -2; msg.img_dim_x_max = 752 + k; msg.img_dim_y_max = 480;
    const int n_kp = 5 + k, n_add = 3 + k;
    const bool sift = (k == 2);                     // one keyframe with CV_32F 128-column descriptors (SIFT branch)
    auto kps = [&](covins::TypeDefs::KeypointVector& kd, covins::TypeDefs::KeypointVector& ku, covins::TypeDefs::AorsVector& ao,
                   cv::Mat& desc, int n, int seed) {
      for (int i = 0; i < n; i++) {
        covins::TypeDefs::KeypointType a, b; fill(a, seed + i); fill(b, seed + 50 + i);
        covins::TypeDefs::AorsType c; fill(c, seed + 100 + i); c(1) = (float)((i + k) % 8);
        kd.push_back(a); ku.push_back(b); ao.push_back(c);
      }
      const int cols = sift ? 128 : 32;
      desc.create(n, cols, sift ? CV_32F : CV_8U);
      for (int i = 0; i < n * cols; i++) {
        if (sift) reinterpret_cast<float*>(desc.ptr())[i] = (float)((i * 7 + seed) % 256);
        else desc.ptr()[i] = (uint8_t)((i * 13 + seed * 5) % 256);
      }
    };
    kps(msg.keypoints_distorted, msg.keypoints_undistorted, msg.keypoints_aors, msg.descriptors, n_kp, 40 + 10 * k);
    kps(msg.keypoints_distorted_add, msg.keypoints_undistorted_add, msg.keypoints_aors_add, msg.descriptors_add, n_add, 70 + 10 * k);
    fill(msg.T_s_c, 1 + k); fill(msg.T_w_s, 2 + k); fill(msg.T_w_s_vio, 3 + k);
    fill(msg.velocity, 4 + k); fill(msg.bias_gyro, 5 + k); fill(msg.bias_accel, 6 + k); fill(msg.lin_acc, 7 + k);
    fill(msg.ang_vel, 8 + k); fill(msg.lin_acc_init, 9 + k); fill(msg.ang_vel_init, 11 + k);
    covins::PreintegrationData& pre = msg.preintegration;
    fill(pre.acc, 12 + k); fill(pre.gyr, 13 + k); fill(pre.lin_bias_accel, 14 + k); fill(pre.lin_bias_gyro, 15 + k);
    for (int s = 0; s < (k == 0 ? 0 : 4 + k); s++) {   // keyframe 0 has no predecessor → empty preintegration
      pre.dt.push_back(0.005); pre.lin_acc_x.push_back(val(s, 1)); pre.lin_acc_y.push_back(val(s, 2)); pre.lin_acc_z.push_back(val(s, 3));
      pre.ang_vel_x.push_back(val(s, 4)); pre.ang_vel_y.push_back(val(s, 5)); pre.ang_vel_z.push_back(val(s, 6));
    }
    for (int i = 0; i < n_kp; i += 2) msg.landmarks.insert(std::make_pair(i, idpair((size_t)(i / 2 + 10 * k), (size_t)kf_ids[k].second)));
    if (k == 1) { msg.id_predecessor = kf_ids[0]; }
    if (k == 0) { msg.id_successor = kf_ids[1]; }
    // msg.img stays the default (empty) cv::Mat, as for keyframes without a stored image
    write_archive(out + "/keyframes/keyframes" + std::to_string(k) + ".txt", msg);
    Json j;
    j.num("timestamp", msg.timestamp); j.pair("id", msg.id);
    j.mat("T_SC", msg.calibration.T_SC); j.num("cam_model", msg.calibration.cam_model); j.num("dist_model", msg.calibration.dist_model);
    j.mat("img_dims", msg.calibration.img_dims); j.mat("dist_coeffs", msg.calibration.dist_coeffs); j.mat("intrinsics", msg.calibration.intrinsics);
    j.mat("K", msg.calibration.K);
    j.num("a_max", msg.calibration.a_max); j.num("g_max", msg.calibration.g_max); j.num("sigma_a_c", msg.calibration.sigma_a_c);
    j.num("sigma_g_c", msg.calibration.sigma_g_c); j.num("sigma_ba", msg.calibration.sigma_ba); j.num("sigma_bg", msg.calibration.sigma_bg);
    j.num("sigma_aw_c", msg.calibration.sigma_aw_c); j.num("sigma_gw_c", msg.calibration.sigma_gw_c); j.num("tau", msg.calibration.tau);
    j.num("g", msg.calibration.g); j.mat("a0", msg.calibration.a0); j.num("rate", msg.calibration.rate);
    j.num("delay_cam0_to_imu", msg.calibration.delay_cam0_to_imu); j.num("delay_cam1_to_imu", msg.calibration.delay_cam1_to_imu);
    j.num("img_dim_x_min", msg.img_dim_x_min); j.num("img_dim_y_min", msg.img_dim_y_min); j.num("img_dim_x_max", msg.img_dim_x_max);
    j.num("img_dim_y_max", msg.img_dim_y_max);
    j.matvec("keypoints_distorted", msg.keypoints_distorted); j.matvec("keypoints_undistorted", msg.keypoints_undistorted);
    j.matvec("keypoints_aors", msg.keypoints_aors); j.cvmat("descriptors", msg.descriptors);
    j.matvec("keypoints_distorted_add", msg.keypoints_distorted_add); j.matvec("keypoints_undistorted_add", msg.keypoints_undistorted_add);
    j.matvec("keypoints_aors_add", msg.keypoints_aors_add); j.cvmat("descriptors_add", msg.descriptors_add);
    j.mat("T_s_c", msg.T_s_c); j.mat("T_w_s", msg.T_w_s); j.mat("T_w_s_vio", msg.T_w_s_vio);
    j.mat("velocity", msg.velocity); j.mat("bias_gyro", msg.bias_gyro); j.mat("bias_accel", msg.bias_accel); j.mat("lin_acc", msg.lin_acc);
    j.mat("ang_vel", msg.ang_vel); j.mat("lin_acc_init", msg.lin_acc_init); j.mat("ang_vel_init", msg.ang_vel_init);
    j.mat("pre_acc", pre.acc); j.mat("pre_gyr", pre.gyr); j.mat("pre_lin_bias_accel", pre.lin_bias_accel); j.mat("pre_lin_bias_gyro", pre.lin_bias_gyro);
    j.vec("pre_dt", pre.dt); j.vec("pre_lin_acc_x", pre.lin_acc_x); j.vec("pre_lin_acc_y", pre.lin_acc_y); j.vec("pre_lin_acc_z", pre.lin_acc_z);
    j.vec("pre_ang_vel_x", pre.ang_vel_x); j.vec("pre_ang_vel_y", pre.ang_vel_y); j.vec("pre_ang_vel_z", pre.ang_vel_z);
    j.key("landmarks"); j.s << "[";
    { bool f = true; for (auto& kv : msg.landmarks) { j.s << (f ? "" : ",") << "[" << kv.first << "," << kv.second.first << "," << kv.second.second << "]"; f = false; } }
    j.s << "]";
    j.pair("id_predecessor", msg.id_predecessor); j.pair("id_successor", msg.id_successor);
    j.cvmat("img", msg.img);
    manifest << (k ? "," : "") << "{" << j.s.str() << "}";
  }
  manifest << "],\"landmarks\":[";
  for (int l = 0; l < 4; l++) {
    covins::MsgLandmark msg;
    msg.save_to_file = true;                       // landmark_be.cpp ConvertToMsgFileExport
    msg.id = idpair((size_t)(3 * l + 1), (size_t)(l % 2));
    fill(msg.pos_w, 60 + l);
    fill(msg.pos_ref, 90 + l);                     // not part of the file-save branch (msg_landmark.hpp:69-73)
    // std::map<idpair,int>: inserted out of key order on purpose, cereal writes in key order
    msg.observations.insert(std::make_pair(kf_ids[2], 4 - l));
    msg.observations.insert(std::make_pair(kf_ids[0], l));
    if (l % 2 == 0) msg.observations.insert(std::make_pair(kf_ids[1], 2 * l));
    msg.id_reference = kf_ids[l % 3];
    write_archive(out + "/mappoints/mappoints" + std::to_string(l) + ".txt", msg);
    Json j;
    j.pair("id", msg.id); j.mat("pos_w", msg.pos_w);
    j.key("observations"); j.s << "[";
    { bool f = true; for (auto& kv : msg.observations) { j.s << (f ? "" : ",") << "[" << kv.first.first << "," << kv.first.second << "," << kv.second << "]"; f = false; } }
    j.s << "]";
    j.pair("id_reference", msg.id_reference);
    manifest << (l ? "," : "") << "{" << j.s.str() << "}";
  }
  manifest << "],\"mapdata\":";
  {
    MsgMapRestated msg;
    msg.id_map = 0;
    msg.keyframes1 = {kf_ids[0], kf_ids[2]};
    msg.keyframes2 = {kf_ids[2], kf_ids[1]};
    for (int e = 0; e < 2; e++) {
      covins::TypeDefs::TransformType T; fill(T, 120 + e);
      covins::TypeDefs::Matrix6Type C; fill(C, 130 + e);
      msg.transforms12.push_back(T); msg.cov.push_back(C);
    }
    write_archive(out + "/mapdata.txt", msg);
    Json j;
    j.num("id_map", (double)msg.id_map);
    j.key("keyframes1"); j.s << "[[0,0],[0,1]]";
    j.key("keyframes2"); j.s << "[[0,1],[1,0]]";
    j.matvec("transforms12", msg.transforms12); j.matvec("cov", msg.cov);
    manifest << "{" << j.s.str() << "}";
  }
  manifest << "}";
  std::ofstream mf(out + "/manifest.json");
  mf << manifest.str();
  return 0;
}

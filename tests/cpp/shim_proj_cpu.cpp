// CPU: covins_b200::SearchByProjection (the shim's host logic: container flattening + replay of the decisions) on mock containers,
// with the C-ABI answered by the oracle-backed test double (stub_cabi_proj.c).  Usage: shim_proj_cpu <dir>
#include <cstdio>
#include <fstream>
#include <string>
#include <vector>

#include "mock_containers.hpp"
#include "../../covins_b200/csrc/host/covins_b200_shim.hpp"

template <class T>
static std::vector<T> rd(const std::string& dir, const char* name) {
  std::ifstream f(dir + "/" + name, std::ios::binary | std::ios::ate);
  if (!f) { std::fprintf(stderr, "missing %s\n", name); std::exit(2); }
  const size_t bytes = (size_t)f.tellg();
  std::vector<T> v(bytes / sizeof(T));
  f.seekg(0);
  f.read(reinterpret_cast<char*>(v.data()), (std::streamsize)bytes);
  return v;
}

int main(int argc, char** argv) {
  if (argc < 2) return 2;
  const std::string dir = argv[1];
  using namespace mock;
  const auto kp = rd<float>(dir, "kf_kp.bin"), oct = rd<float>(dir, "kf_octave.bin");
  const auto kdesc = rd<unsigned char>(dir, "kf_desc.bin"), has_lm = rd<unsigned char>(dir, "kf_has_lm.bin"), matched = rd<unsigned char>(dir, "matched.bin");
  const auto cand = rd<int32_t>(dir, "kf_lm_cand.bin"), feat = rd<int32_t>(dir, "lm_feat_idx.bin");
  const auto tcw = rd<double>(dir, "Tcw.bin"), intr = rd<double>(dir, "intr.bin"), dist = rd<double>(dir, "dist.bin");
  const auto lvalid = rd<unsigned char>(dir, "lm_valid.bin"), ldesc = rd<unsigned char>(dir, "lm_desc.bin");
  const auto lpos = rd<double>(dir, "lm_pos.bin"), lnormal = rd<double>(dir, "lm_normal.bin"), lmind = rd<double>(dir, "lm_min_distance.bin"),
             lmaxd = rd<double>(dir, "lm_max_distance.bin");
  const size_t n = oct.size(), m = lvalid.size();
  auto kf = std::make_shared<Keyframe>();
  kf->keypoints_distorted_.resize(n); kf->keypoints_aors_.resize(n); kf->descriptors_.resize(n); kf->landmarks_.assign(n, nullptr);
  for (size_t i = 0; i < n; i++) {
    kf->keypoints_distorted_[i] = {kp[2 * i], kp[2 * i + 1]};
    kf->keypoints_aors_[i] = {0.f, oct[i], 0.f, 0.f};
    std::copy(kdesc.begin() + 32 * i, kdesc.begin() + 32 * i + 32, kf->descriptors_[i].begin());
  }
  for (int k = 0; k < 4; k++) { kf->intr[k] = intr[k]; kf->dist[k] = dist[k]; }
  std::vector<LandmarkPtr> pts(m);
  for (size_t i = 0; i < m; i++) {
    auto lm = std::make_shared<Landmark>();
    lm->invalid = !lvalid[i];
    lm->pos_w_ = {lpos[3 * i], lpos[3 * i + 1], lpos[3 * i + 2]};
    lm->normal_ = {lnormal[3 * i], lnormal[3 * i + 1], lnormal[3 * i + 2]};
    lm->min_distance_ = lmind[i]; lm->max_distance_ = lmaxd[i];
    std::copy(ldesc.begin() + 32 * i, ldesc.begin() + 32 * i + 32, lm->descriptor_.begin());
    if (feat[i] >= 0) lm->observations_[kf] = (size_t)feat[i];
    pts[i] = lm;
  }
  auto other = std::make_shared<Landmark>();     // a landmark that is not in the candidate list
  for (size_t i = 0; i < n; i++)
    if (has_lm[i]) kf->landmarks_[i] = cand[i] >= 0 ? pts[cand[i]] : other;
  std::vector<LandmarkPtr> vpMatched(n, nullptr);
  for (size_t i = 0; i < n; i++)
    if (matched[i]) vpMatched[i] = other;
  Transform Tcw;
  for (int r = 0; r < 4; r++) for (int c = 0; c < 4; c++) Tcw(r, c) = tcw[4 * r + c];
  covins_b200::Context ctx(0);
  const int nm = covins_b200::SearchByProjection(ctx, kf, Tcw, pts, vpMatched, 10.0, 50, 1, 2.0);
  std::map<const Landmark*, int> idx;
  for (size_t i = 0; i < m; i++) idx[pts[i].get()] = (int)i;
  auto code = [&](const LandmarkPtr& p) -> int32_t { return !p ? -1 : (p == other ? -2 : idx.at(p.get())); };
  std::vector<int32_t> out{nm};
  for (size_t i = 0; i < n; i++) out.push_back(code(vpMatched[i]));
  for (size_t i = 0; i < n; i++) out.push_back(code(kf->landmarks_[i]));
  for (size_t i = 0; i < m; i++) out.push_back(pts[i]->GetFeatureIndex(kf));
  std::ofstream(dir + "/proj_out.bin", std::ios::binary).write(reinterpret_cast<const char*>(out.data()), (std::streamsize)(out.size() * 4));
  std::printf("nmatches %d\n", nm);
  return 0;
}

// oracle/ref/dm_ref_shim.cpp — TEST INFRASTRUCTURE ONLY (checker; never linked into covins_b200/).
//
// Drives the REFERENCE's own estd2::DenseMatcher (compiled from the sources where they lie under
// /root/reference/covins_backend: src/dense_matcher/DenseMatcher.cpp, src/dense_matcher/ThreadPool.cpp,
// src/matcher/MatchingAlgorithm.cpp + include/covins/dense_matcher/implementation/DenseMatcher.hpp) through a
// byte-array MatchingAlgorithm, so that the matching stages M5/M6 have REFERENCE-PRODUCED results to
// pin oracle/knn_oracle.c:ora_landmark_match and the CUDA path against.  Nothing of the reference is copied: the
// Makefile next to this file compiles the reference translation units in place into oracle/_ref/libdm_ref.so.
//
// ByteArrayMatchingAlgorithm is written here (the reference's LandmarkMatchingAlgorithm needs Keyframe/Eigen/OpenCV):
//   distance()   = 256-bit Hamming as 8 x int32 popcounts (what FeatureMatcher::DescriptorDistanceHamming computes,
//                  feature_matcher_be.cpp:49-64), FLT_MAX unless dist < threshold (LandmarkMatchingAlgorithm.h:103-114;
//                  verifyMatch is constant true, LandmarkMatchingAlgorithm.cpp:122-135)
//   skipA/skipB  = the caller's masks (LandmarkMatchingAlgorithm.cpp:76-84, 93-101)
//   setBestMatch = append Match(idxA, idxB, distance) (LandmarkMatchingAlgorithm.cpp:155-163)
// DenseMatcher(numThreads, numBest=4, useDistanceRatioThreshold=false): placerec_be.cpp:85-90 passes 8 threads;
// the canonical (deterministic) order of SURVEY §8c is numThreads = 1.
#include <cstdint>
#include <cstring>
#include <limits>
#include <vector>

#include "covins/dense_matcher/DenseMatcher.hpp"
#include "matcher/MatchingAlgorithm.h"

namespace {

class ByteArrayMatchingAlgorithm : public covins::MatchingAlgorithm {
 public:
  ByteArrayMatchingAlgorithm(const uint8_t* A, const uint8_t* skipA, size_t nA, const uint8_t* B, const uint8_t* skipB,
                             size_t nB, float thr)
      : A_(A), skipA_(skipA), nA_(nA), B_(B), skipB_(skipB), nB_(nB), thr_(thr) {}
  size_t sizeA() const override { return nA_; }
  size_t sizeB() const override { return nB_; }
  float distanceThreshold() const override { return thr_; }
  bool skipA(size_t i) const override { return skipA_ && skipA_[i]; }
  bool skipB(size_t i) const override { return skipB_ && skipB_[i]; }
  float distance(size_t ia, size_t ib) const override {
    uint32_t a[8], b[8];
    std::memcpy(a, A_ + 32 * ia, 32);
    std::memcpy(b, B_ + 32 * ib, 32);
    int d = 0;
    for (int i = 0; i < 8; i++) d += __builtin_popcount(a[i] ^ b[i]);
    const float dist = static_cast<float>(d);
    return dist < thr_ ? dist : std::numeric_limits<float>::max();
  }
  void reserveMatches(size_t) override {}
  void setBestMatch(size_t ia, size_t ib, double d) override { matches.emplace_back(ia, ib, (float)d); }
  covins::Matches matches;

 private:
  const uint8_t *A_, *skipA_;
  size_t nA_;
  const uint8_t *B_, *skipB_;
  size_t nB_;
  float thr_;
};

}  // namespace

// returns the number of matches (ordered by B index, as matchBody emits them), at most nB
extern "C" __attribute__((visibility("default"))) int dm_ref_match(const uint8_t* A, const uint8_t* skipA, int nA,
                                                                   const uint8_t* B, const uint8_t* skipB, int nB,
                                                                   float thr, int num_threads, int num_best,
                                                                   int32_t* outA, int32_t* outB, float* outD) {
  ByteArrayMatchingAlgorithm algo(A, skipA, (size_t)nA, B, skipB, (size_t)nB, thr);
  estd2::DenseMatcher matcher((unsigned char)num_threads, (unsigned char)num_best, false);
  matcher.match<ByteArrayMatchingAlgorithm>(algo);
  int n = 0;
  for (const auto& m : algo.matches) {
    outA[n] = (int32_t)m.idxA;
    outB[n] = (int32_t)m.idxB;
    outD[n] = m.distance;
    n++;
  }
  return n;
}

/*
 * oracle/knn_oracle.c — CPU restatement of the COVINS place-recognition matching stage.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under covins_b200/ may include, link or call this file;
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs use it
 * (as the checker / the timed CPU baseline, never as the product path).
 *
 * What it restates (all paths relative to /root/reference/covins_backend):
 *   ora_knn_hamming      cv::BFMatcher(NORM_HAMMING)::knnMatch(q, t, out, k)
 *                        call sites src/covins_backend/placerec_gen_be.cpp:82-100,
 *                        src/covins_backend/RelNonCentralPosSolver.cpp:303-324.
 *                        OpenCV is a third-party dependency that is NOT in the tree
 *                        (dependencies.rosinstall:31-33, opencv3_catkin, unpinned).  Its published
 *                        algorithm (modules/core/src/batch_distance.cpp, batchDistance with K>0):
 *                        per query row scan train rows ascending; a candidate enters the K-list
 *                        only if d < dist[K-1] (strict); it is inserted AFTER all entries with
 *                        dist <= d, so equal distances keep ascending trainIdx.
 *                        PINNED: tests/golden/knn_*.npz were produced by cv2 4.13 BFMatcher in the
 *                        build container (tests/golden/gen_golden.py); tests/test_oracle_knn.py
 *                        checks this file against them bit-for-bit.
 *   ora_knn_l2           cv::BFMatcher(NORM_L2)::knnMatch on CV_32F rows: d = sqrtf(sum (a-b)^2),
 *                        fp32 accumulation.  For integer-valued descriptors in [0,255] (what
 *                        cv::xfeatures2d::SIFT produces, covins_frontend/src/frontend_wrapper.cpp:603)
 *                        every partial sum is an integer < 2^24, so the result is independent of
 *                        summation order and bit-exact against cv2 (golden pinned).  The reference
 *                        uses cv::FlannBasedMatcher for SIFT (placerec_gen_be.cpp:86-87), an
 *                        approximate randomised kd-forest that is not reproducible; the exact
 *                        brute-force result is the parity target.
 *   ora_ratio_filter     placerec_gen_be.cpp:102-114 == RelNonCentralPosSolver.cpp:326-337
 *                        keep m iff m.distance <= img_match_thres && m.distance < ratio_thres*n.distance
 *                        (all three are float: include/covins/covins_base/config_backend.hpp:119-120).
 *   ora_hamming256       FeatureMatcher::DescriptorDistanceHamming, src/covins_backend/feature_matcher_be.cpp:49-64
 *                        (8 x u32 SWAR popcount, hard-coded 256 bit).
 *   ora_landmark_descriptor  Landmark::ComputeDescriptor, src/covins_backend/landmark_be.cpp:49-92 (representative
 *                        descriptor = the observation with the least median Hamming distance to all observations).
 *   ora_landmark_match   estd2::DenseMatcher::match<LandmarkMatchingAlgorithm> as used at
 *                        src/covins_backend/placerec_be.cpp:85-90:
 *                          distance()            include/covins/matcher/LandmarkMatchingAlgorithm.h:103-114
 *                          doWorkLinearMatching  include/covins/dense_matcher/implementation/DenseMatcher.hpp:180-220
 *                          listBIteration        .../implementation/DenseMatcher.hpp:152-176
 *                          assignbest            src/dense_matcher/DenseMatcher.cpp:62-104
 *                          final sweep           .../implementation/DenseMatcher.hpp:93-121
 *                        These files are fully in-tree, so this restatement is exact — with ONE
 *                        canonicalisation: the reference runs 8 threads striding A and its tie
 *                        outcome depends on thread arrival order (first-come wins on equal
 *                        distance).  The oracle processes A = 0,1,2,... sequentially, i.e. the
 *                        reference with numMatcherThreads = 1.  No reference test pins this stage
 *                        (the reference has no tests): "parity unpinned" beyond the in-tree source.
 *
 * Build: see oracle/Makefile (gcc -O3 -march=native -fopenmp -shared -fPIC).
 */
#include <float.h>
#include <limits.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORA_API __attribute__((visibility("default")))

/* feature_matcher_be.cpp:49-64 — bit-hack popcount over 8 x int32. */
ORA_API int ora_hamming256(const uint8_t *a, const uint8_t *b) {
  const uint32_t *pa = (const uint32_t *)a;
  const uint32_t *pb = (const uint32_t *)b;
  int dist = 0;
  for (int i = 0; i < 8; i++, pa++, pb++) {
    uint32_t v = *pa ^ *pb;
    v = v - ((v >> 1) & 0x55555555u);
    v = (v & 0x33333333u) + ((v >> 2) & 0x33333333u);
    dist += (((v + (v >> 4)) & 0xF0F0F0Fu) * 0x1010101u) >> 24;
  }
  return dist;
}

/* Generic byte-length Hamming (cv::norm NORM_HAMMING over `bytes` bytes). */
static inline int hamming_bytes(const uint8_t *a, const uint8_t *b, int bytes) {
  int d = 0, i = 0;
  for (; i + 8 <= bytes; i += 8) {
    uint64_t x, y;
    memcpy(&x, a + i, 8);
    memcpy(&y, b + i, 8);
    d += __builtin_popcountll(x ^ y);
  }
  for (; i < bytes; i++) d += __builtin_popcount((unsigned)(a[i] ^ b[i]));
  return d;
}

/* batchDistance K-list update: strict '<' vs current worst, insert after equals. */
static inline void klist_insert_i(int32_t *idx, int32_t *dist, int k, int32_t j, int32_t d) {
  if (!(d < dist[k - 1])) return;
  int p = k - 1;
  while (p > 0 && dist[p - 1] > d) { /* move strictly-greater entries down */
    dist[p] = dist[p - 1];
    idx[p] = idx[p - 1];
    p--;
  }
  dist[p] = d;
  idx[p] = j;
}
static inline void klist_insert_f(int32_t *idx, float *dist, int k, int32_t j, float d) {
  if (!(d < dist[k - 1])) return;
  int p = k - 1;
  while (p > 0 && dist[p - 1] > d) {
    dist[p] = dist[p - 1];
    idx[p] = idx[p - 1];
    p--;
  }
  dist[p] = d;
  idx[p] = j;
}

/*
 * BFMatcher(NORM_HAMMING).knnMatch.  idx/dist are [nq][k]; unused slots (nt < k) hold idx -1,
 * dist INT32_MAX (OpenCV returns shorter inner vectors there).  bytes = descriptor length (32 ORB).
 * threads <= 0 → all OpenMP threads.
 */
ORA_API void ora_knn_hamming(const uint8_t *q, int nq, const uint8_t *t, int nt, int bytes, int k,
                             int32_t *idx, int32_t *dist, int threads) {
#ifdef _OPENMP
  if (threads > 0) omp_set_num_threads(threads);
#endif
#pragma omp parallel for schedule(static)
  for (int i = 0; i < nq; i++) {
    int32_t *ii = idx + (size_t)i * k, *dd = dist + (size_t)i * k;
    for (int s = 0; s < k; s++) { ii[s] = -1; dd[s] = INT32_MAX; }
    const uint8_t *qi = q + (size_t)i * bytes;
    for (int j = 0; j < nt; j++) {
      int d = hamming_bytes(qi, t + (size_t)j * bytes, bytes);
      klist_insert_i(ii, dd, k, j, d);
    }
  }
}

/* Segmented form: train set = concatenation of n_seg candidate keyframes, seg_ptr[n_seg+1] row
 * offsets; one independent knnMatch per (segment, query) exactly like the per-candidate loop at
 * placerec_gen_be.cpp:72-125.  Output [n_seg][nq][k], trainIdx LOCAL to the segment. */
ORA_API void ora_knn_hamming_batch(const uint8_t *q, int nq, const uint8_t *t, const int32_t *seg_ptr,
                                   int n_seg, int bytes, int k, int32_t *idx, int32_t *dist,
                                   int threads) {
#ifdef _OPENMP
  if (threads > 0) omp_set_num_threads(threads);
#endif
  const int QB = 32; /* query block per task: (segment x query block) tasks keep every core busy */
  const int nqb = (nq + QB - 1) / QB;
#pragma omp parallel for collapse(2) schedule(dynamic, 4)
  for (int s = 0; s < n_seg; s++) {
    for (int b = 0; b < nqb; b++) {
      const uint8_t *ts = t + (size_t)seg_ptr[s] * bytes;
      int nt = seg_ptr[s + 1] - seg_ptr[s];
      int i1 = (b + 1) * QB < nq ? (b + 1) * QB : nq;
      for (int i = b * QB; i < i1; i++) {
        int32_t *ii = idx + ((size_t)s * nq + i) * k, *dd = dist + ((size_t)s * nq + i) * k;
        for (int c = 0; c < k; c++) { ii[c] = -1; dd[c] = INT32_MAX; }
        const uint8_t *qi = q + (size_t)i * bytes;
        for (int j = 0; j < nt; j++) {
          int d = hamming_bytes(qi, ts + (size_t)j * bytes, bytes);
          klist_insert_i(ii, dd, k, j, d);
        }
      }
    }
  }
}

/* BFMatcher(NORM_L2).knnMatch on CV_32F: fp32 accumulate of (a-b)^2, then sqrtf. */
ORA_API void ora_knn_l2_batch(const float *q, int nq, const float *t, const int32_t *seg_ptr, int n_seg,
                              int dim, int k, int32_t *idx, float *dist, int threads) {
#ifdef _OPENMP
  if (threads > 0) omp_set_num_threads(threads);
#endif
#pragma omp parallel for schedule(dynamic, 1)
  for (int s = 0; s < n_seg; s++) {
    const float *ts = t + (size_t)seg_ptr[s] * dim;
    int nt = seg_ptr[s + 1] - seg_ptr[s];
    for (int i = 0; i < nq; i++) {
      int32_t *ii = idx + ((size_t)s * nq + i) * k;
      float *dd = dist + ((size_t)s * nq + i) * k;
      for (int c = 0; c < k; c++) { ii[c] = -1; dd[c] = FLT_MAX; }
      const float *qi = q + (size_t)i * dim;
      for (int j = 0; j < nt; j++) {
        const float *tj = ts + (size_t)j * dim;
        float acc = 0.f;
        for (int c = 0; c < dim; c++) {
          float df = qi[c] - tj[c];
          acc += df * df;
        }
        klist_insert_f(ii, dd, k, j, sqrtf(acc));
      }
    }
  }
}

/*
 * placerec_gen_be.cpp:102-114.  Inputs are the k=2 lists ([n][2]) as float distances.  Writes, per
 * query row, the accepted trainIdx or -1; returns the number of accepted matches (== the
 * reference's img_matches.size(), compared with matches_thres at :116-124).  Rows whose second
 * neighbour is missing (idx -1) are rejected — the reference would index out of range there.
 */
ORA_API int ora_ratio_filter(const int32_t *idx2, const float *dist2, int n, float thr, float ratio,
                             int32_t *match_train, float *match_dist) {
  int cnt = 0;
  for (int i = 0; i < n; i++) {
    int32_t m = idx2[2 * i], nn = idx2[2 * i + 1];
    float dm = dist2[2 * i], dn = dist2[2 * i + 1];
    int ok = 0;
    if (m >= 0 && nn >= 0) {
      if (dm <= thr) {
        volatile float rhs = ratio * dn; /* float*float rounded to float, as in the reference */
        if (dm < rhs) ok = 1;
      }
    }
    match_train[i] = ok ? m : -1;
    match_dist[i] = ok ? dm : FLT_MAX;
    cnt += ok;
  }
  return cnt;
}

/* ---------------------------------------------------------------------------------------------
 * DenseMatcher<LandmarkMatchingAlgorithm> (COVINS mode), canonical A-sequential order.
 * ------------------------------------------------------------------------------------------- */
typedef struct { int indexA; float distance; } pairing_t; /* DenseMatcher.hpp:89-113 */

/* LandmarkMatchingAlgorithm.h:103-114: float(dist) if dist < threshold else FLT_MAX
 * (verifyMatch is hard-wired to true, LandmarkMatchingAlgorithm.cpp:123-137). */
static inline float lm_distance(const uint8_t *a, const uint8_t *b, float thr) {
  float d = (float)ora_hamming256(a, b);
  return (d < thr) ? d : FLT_MAX;
}

/* DenseMatcher.cpp:62-104 without the mutexes (single thread). Iterative form of the recursion. */
static void assignbest(int a, pairing_t *vpairs, const pairing_t *best, int numBest, int startidx) {
  for (;;) {
    const pairing_t *ai = best + (size_t)a * numBest;
    int reassigned = 0;
    for (int index = startidx; index < numBest && ai[index].indexA != -1; ++index) {
      int b = ai[index].indexA;
      if (vpairs[b].indexA == -1) {
        vpairs[b].indexA = a;
        vpairs[b].distance = ai[index].distance;
        return;
      } else if (ai[index].distance < vpairs[b].distance) {
        int old = vpairs[b].indexA;
        vpairs[b].indexA = a;
        vpairs[b].distance = ai[index].distance;
        a = old;       /* reassign the displaced A ... */
        startidx = 1;  /* ... from its list position 1 (DenseMatcher.cpp:97-98) */
        reassigned = 1;
        break;
      }
    }
    if (!reassigned) return;
  }
}

/*
 * A: [nA][32] descriptors of the query KF, skipA[nA] (1 = keypoint has no valid landmark,
 * LandmarkMatchingAlgorithm.cpp:76-84), B likewise for the candidate KF.  numBest = 4, thr = 50.0
 * in the reference (DenseMatcher.hpp:68, placerec_be.cpp:86).  Outputs matches ordered by B index:
 * (outA[m], outB[m], outD[m]); returns the count.  best_out (nullable): [nA][numBest] (idxB, dist)
 * lists for debugging / kernel parity, idxB = -1 for empty slots.
 */
ORA_API int ora_landmark_match(const uint8_t *A, const uint8_t *skipA, int nA, const uint8_t *B,
                               const uint8_t *skipB, int nB, float thr, int numBest, int32_t *outA,
                               int32_t *outB, float *outD, int32_t *best_idx_out, float *best_dist_out) {
  pairing_t *best = (pairing_t *)malloc(sizeof(pairing_t) * (size_t)nA * numBest);
  pairing_t *vpairs = (pairing_t *)malloc(sizeof(pairing_t) * (size_t)(nB > 0 ? nB : 1));
  for (int b = 0; b < nB; b++) { vpairs[b].indexA = -1; vpairs[b].distance = FLT_MAX; }
  for (size_t i = 0; i < (size_t)nA * numBest; i++) { best[i].indexA = -1; best[i].distance = thr; }

  for (int a = 0; a < nA; a++) {
    if (skipA && skipA[a]) continue; /* DenseMatcher.hpp:194-195 */
    pairing_t *ai = best + (size_t)a * numBest;
    for (int b = 0; b < nB; b++) {
      if (skipB && skipB[b]) continue;
      float d = lm_distance(A + (size_t)a * 32, B + (size_t)b * 32, thr);
      if (d < ai[numBest - 1].distance) { /* strict (:159) */
        /* std::lower_bound on distance: first position with distance >= d → before equals */
        int lb = 0;
        while (lb < numBest && ai[lb].distance < d) lb++;
        for (int p = numBest - 1; p > lb; p--) ai[p] = ai[p - 1];
        ai[lb].indexA = b;
        ai[lb].distance = d;
      }
    }
    assignbest(a, vpairs, best, numBest, 0); /* DenseMatcher.hpp:213-214 */
  }
  int cnt = 0;
  for (int b = 0; b < nB; b++) { /* final sweep, ratio branch off (useDistanceRatioThreshold_=false) */
    if (vpairs[b].distance < thr) {
      outA[cnt] = vpairs[b].indexA;
      outB[cnt] = b;
      outD[cnt] = vpairs[b].distance;
      cnt++;
    }
  }
  if (best_idx_out && best_dist_out)
    for (size_t i = 0; i < (size_t)nA * numBest; i++) {
      best_idx_out[i] = best[i].indexA;
      best_dist_out[i] = best[i].distance;
    }
  free(best);
  free(vpairs);
  return cnt;
}

/* Batched over candidates, 8 threads over candidates for the timed CPU baseline. */
ORA_API void ora_landmark_match_batch(const uint8_t *A, const uint8_t *skipA, int nA, const uint8_t *B,
                                      const uint8_t *skipB, const int32_t *seg_ptr, int n_seg, float thr,
                                      int numBest, int32_t *outA, int32_t *outB, float *outD,
                                      int32_t *n_out, int threads) {
#ifdef _OPENMP
  if (threads > 0) omp_set_num_threads(threads);
#endif
#pragma omp parallel for schedule(dynamic, 1)
  for (int s = 0; s < n_seg; s++) {
    int off = seg_ptr[s], nB = seg_ptr[s + 1] - seg_ptr[s];
    n_out[s] = ora_landmark_match(A, skipA, nA, B + (size_t)off * 32, skipB ? skipB + off : NULL, nB, thr,
                                  numBest, outA + off, outB + off, outD + off, NULL, NULL);
  }
}


/* ------------------------------------------------------------------------------------------------
 * Landmark::ComputeDescriptor (src/covins_backend/landmark_be.cpp:49-92), batched over landmarks.
 * cand: the descriptor rows (32 B) of the valid observing keyframes of every landmark, concatenated in
 * the landmark's observation order (:57-64; canonical order = keyframe idpair order);
 * lm_ptr[n_lm+1] row offsets.  Per landmark: the full num_desc x num_desc Hamming matrix with a zero
 * diagonal (:70-78), per row the sorted distances and median = sorted[(int)(0.5 * (num_desc - 1))]
 * (:84-85), the FIRST row with the strictly smallest median wins (:86-89).  best_idx[l] = that row
 * (landmark-local) or -1 for a landmark without candidates (the reference returns early and keeps the
 * old descriptor, :53-55,65-67); out_desc[l] = its 32 bytes (untouched when -1).
 * ------------------------------------------------------------------------------------------------ */
static int cmp_double(const void *a, const void *b) {
  double x = *(const double *)a, y = *(const double *)b;
  return (x > y) - (x < y);
}
ORA_API void ora_landmark_descriptor(const uint8_t *cand, const int32_t *lm_ptr, int n_lm, int32_t *best_idx,
                                     uint8_t *out_desc) {
  for (int l = 0; l < n_lm; l++) {
    const int n = lm_ptr[l + 1] - lm_ptr[l];
    best_idx[l] = -1;
    if (n <= 0) continue;
    const uint8_t *D = cand + (size_t)lm_ptr[l] * 32;
    double *dist = (double *)malloc(sizeof(double) * (size_t)n * n);
    double *row = (double *)malloc(sizeof(double) * (size_t)n);
    for (int i = 0; i < n; i++) {
      dist[(size_t)i * n + i] = 0;
      for (int j = i + 1; j < n; j++) {
        int d = 0;
        for (int b = 0; b < 32; b++) d += __builtin_popcount((unsigned)(D[(size_t)i * 32 + b] ^ D[(size_t)j * 32 + b]));
        dist[(size_t)i * n + j] = dist[(size_t)j * n + i] = (double)d;
      }
    }
    double best_median = (double)INT_MAX;
    int best = -1;
    for (int i = 0; i < n; i++) {
      memcpy(row, dist + (size_t)i * n, sizeof(double) * (size_t)n);
      qsort(row, (size_t)n, sizeof(double), cmp_double);
      const double median = row[(int)(0.5 * (n - 1))];
      if (median < best_median) { best_median = median; best = i; }
    }
    best_idx[l] = best;
    memcpy(out_desc + (size_t)l * 32, D + (size_t)best * 32, 32);
    free(dist); free(row);
  }
}

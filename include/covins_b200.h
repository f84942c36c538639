/*
 * covins_b200.h — C-ABI of libcovins_b200.so: the B200-native (sm_100a) implementation of the COVINS
 * server hot path (place-recognition descriptor matching + PGO / global BA).
 *
 * This is the drop-in boundary.  The reference has no FFI layer — its seams are C++
 * symbols — so each entry point below names the reference call it replaces (paths relative to
 * covins_backend/ in VIS4ROB-lab/covins).  The C++ host shim that keeps the reference's own signatures
 * (Optimization::*, the ComputeSE3 matching blocks) and calls these functions is
 * covins_b200/csrc/host/covins_b200_shim.hpp; the binding a maintainer adds is shown in INTEGRATION.md.
 *
 * Conventions
 *   - plain pointers + sizes, no C++/torch types; every function returns a cvb_status (0 = OK);
 *     cvb_last_error(ctx) gives the message of the last failure on that ctx.
 *   - functions without suffix take HOST buffers and include all H2D/D2H copies (synchronous on
 *     return); `_dev` variants take DEVICE pointers and enqueue on `stream` (a cudaStream_t passed as
 *     void*; NULL = the ctx's own stream) without synchronising.
 *   - a ctx owns one device, one stream, and grow-only device workspaces; it is NOT thread-safe:
 *     use one ctx per host thread (the reference runs one place-recognition thread per agent,
 *     src/covins_backend/handler_be.cpp:52-56, and at most one optimisation per map).
 *   - no CPU fallback exists: without a CUDA device every compute call fails with CVB_ERR_CUDA.
 */
#ifndef COVINS_B200_H_
#define COVINS_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CVB_API __attribute__((visibility("default")))

typedef enum cvb_status {
  CVB_OK = 0,
  CVB_ERR_INVALID = 1,     /* bad argument */
  CVB_ERR_CUDA = 2,        /* CUDA runtime error / no device */
  CVB_ERR_UNSUPPORTED = 3, /* valid request outside the implemented envelope */
  CVB_ERR_NUMERIC = 4      /* factorisation failed / non-finite state */
} cvb_status;

typedef struct cvb_ctx cvb_ctx;

CVB_API int cvb_version(void);
CVB_API int cvb_ctx_create(int device, cvb_ctx** out);
CVB_API int cvb_ctx_destroy(cvb_ctx* ctx);
CVB_API const char* cvb_last_error(const cvb_ctx* ctx);
CVB_API int cvb_ctx_sync(cvb_ctx* ctx);
/* number of kernels this ctx has launched so far (bench.py's gpu_launches) */
CVB_API int64_t cvb_launch_count(const cvb_ctx* ctx);

/* ------------------------------------------------------------------------------------------------
 * Matching half
 * ---------------------------------------------------------------------------------------------- */

/*
 * Replaces cv::BFMatcher(cv::NORM_HAMMING)::knnMatch(query, train, out, k)
 *   src/covins_backend/placerec_gen_be.cpp:82-100, src/covins_backend/RelNonCentralPosSolver.cpp:303-324.
 * The train set is the concatenation of n_seg candidate keyframes (seg_ptr[n_seg+1] row offsets,
 * seg_ptr[0] = 0); one independent knnMatch per (segment, query), exactly the per-candidate loop of
 * placerec_gen_be.cpp:72-125 — n_seg = 1 is the plain call.  Descriptors are 32-byte ORB rows
 * (desc_length 32, config/config_backend.yaml:28-29).
 * Output idx/dist are [n_seg][nq][k] (k in 1..4): trainIdx LOCAL to the segment, Hamming distance as
 * int32 (DMatch::distance is this value as float).  Order and ties as OpenCV: ascending distance,
 * equal distances by ascending trainIdx.  Slots beyond the segment length: idx -1, dist INT32_MAX.
 */
CVB_API int cvb_knn_hamming_batch(cvb_ctx* ctx, const uint8_t* q, int nq, const uint8_t* t,
                                  const int32_t* seg_ptr, int n_seg, int k, int32_t* idx, int32_t* dist);
CVB_API int cvb_knn_hamming_batch_dev(cvb_ctx* ctx, const uint8_t* d_q, int nq, const uint8_t* d_t,
                                      const int32_t* d_seg_ptr, const int32_t* h_seg_ptr, int n_seg, int k,
                                      int32_t* d_idx, int32_t* d_dist, void* stream);

/*
 * knnMatch(k=2) fused with the distance + Lowe-ratio filter of placerec_gen_be.cpp:102-114
 * (== RelNonCentralPosSolver.cpp:326-337): keep m iff m.distance <= thr && m.distance < ratio*n.distance,
 * float arithmetic as in the reference (config_backend.hpp:119-120 reads both as float).
 * match_train [n_seg][nq]: accepted trainIdx (segment-local) or -1; match_dist [n_seg][nq]: its
 * distance (float) or FLT_MAX; n_matches [n_seg]: img_matches.size(), the number compared with
 * matches_thres at placerec_gen_be.cpp:116-124.  The reference's Matches vector is the accepted rows
 * in query order.
 */
CVB_API int cvb_match_hamming_batch(cvb_ctx* ctx, const uint8_t* q, int nq, const uint8_t* t,
                                    const int32_t* seg_ptr, int n_seg, float thr, float ratio,
                                    int32_t* match_train, float* match_dist, int32_t* n_matches);
CVB_API int cvb_match_hamming_batch_dev(cvb_ctx* ctx, const uint8_t* d_q, int nq, const uint8_t* d_t,
                                        const int32_t* d_seg_ptr, const int32_t* h_seg_ptr, int n_seg,
                                        float thr, float ratio, int32_t* d_match_train, float* d_match_dist,
                                        int32_t* d_n_matches, void* stream);

/*
 * Replaces the SIFT branch, cv::FlannBasedMatcher()::knnMatch(query, train, out, 2)
 *   (placerec_gen_be.cpp:86-87,99; RelNonCentralPosSolver.cpp:310-311,323), with the EXACT brute-force
 *   result cv::BFMatcher(NORM_L2) gives (FLANN is approximate and randomised).
 * q/t are CV_32F rows of `dim` floats (dim % 16 == 0, dim <= 256; SIFT: 128).  Descriptors must be
 * integer-valued in [0,255] (what cv::xfeatures2d::SIFT emits) — then every fp32 partial sum of the
 * reference is exact and the result is bit-identical to OpenCV; other inputs → CVB_ERR_UNSUPPORTED.
 * dist = sqrtf(sum (a-b)^2) as float.
 */
CVB_API int cvb_knn_l2_batch(cvb_ctx* ctx, const float* q, int nq, const float* t, const int32_t* seg_ptr,
                             int n_seg, int dim, int k, int32_t* idx, float* dist);
/* device variant on the HBM-resident layout: descriptors quantised to u8 [rows][dim] (exact) */
CVB_API int cvb_knn_l2_u8_batch_dev(cvb_ctx* ctx, const uint8_t* d_q, int nq, const uint8_t* d_t,
                                    const int32_t* d_seg_ptr, const int32_t* h_seg_ptr, int n_seg, int dim,
                                    int k, int32_t* d_idx, float* d_dist, void* stream);
CVB_API int cvb_match_l2_batch(cvb_ctx* ctx, const float* q, int nq, const float* t, const int32_t* seg_ptr,
                               int n_seg, int dim, float thr, float ratio, int32_t* match_train,
                               float* match_dist, int32_t* n_matches);
/* f32 [rows][dim] (device) → u8 [rows][dim] (device); *d_bad (int32, device) is set to 1 if any value is
 * not an integer in [0,255]. */
CVB_API int cvb_quantize_u8_dev(cvb_ctx* ctx, const float* d_src, int64_t n, uint8_t* d_dst, int32_t* d_bad,
                                void* stream);

/*
 * Replaces estd2::DenseMatcher(8).match<LandmarkMatchingAlgorithm>(algo) with
 * LandmarkMatchingAlgorithm(50.0), src/covins_backend/placerec_be.cpp:85-90:
 *   top-numBest scan   include/covins/dense_matcher/implementation/DenseMatcher.hpp:152-220
 *   mutual assignment  src/dense_matcher/DenseMatcher.cpp:62-104
 *   final sweep        include/covins/dense_matcher/implementation/DenseMatcher.hpp:93-121
 *   distance           include/covins/matcher/LandmarkMatchingAlgorithm.h:103-114 (256-bit Hamming,
 *                      src/covins_backend/feature_matcher_be.cpp:49-64; FLT_MAX if >= thr)
 * in the canonical A-sequential order (the reference with numMatcherThreads = 1; with 8 threads its
 * tie outcome depends on thread arrival order).  A = query KF descriptors [nA][32], skipA[nA] = 1 where
 * the keypoint has no valid landmark (LandmarkMatchingAlgorithm.cpp:76-84); B = concatenated candidate
 * KFs with seg_ptr as above.  num_best in 1..4 (reference: 4), thr 50.0.
 * Output per segment s, in the slice [seg_ptr[s], seg_ptr[s]+n_out[s]): matches ordered by B index,
 * outA = idxA, outB = idxB (segment-local), outD = distance — the Matches vector of placerec_be.cpp:91.
 */
CVB_API int cvb_landmark_match_batch(cvb_ctx* ctx, const uint8_t* A, const uint8_t* skipA, int nA,
                                     const uint8_t* B, const uint8_t* skipB, const int32_t* seg_ptr,
                                     int n_seg, float thr, int num_best, int32_t* outA, int32_t* outB,
                                     float* outD, int32_t* n_out);
CVB_API int cvb_landmark_match_batch_dev(cvb_ctx* ctx, const uint8_t* d_A, const uint8_t* d_skipA, int nA,
                                         const uint8_t* d_B, const uint8_t* d_skipB, const int32_t* d_seg_ptr,
                                         const int32_t* h_seg_ptr, int n_seg, float thr, int num_best,
                                         int32_t* d_outA, int32_t* d_outB, float* d_outD, int32_t* d_n_out,
                                         void* stream);

/* INT-pipe microbenchmark used for the Hamming roofline denominator (SURVEY.md §8d asks the builder to
 * measure the popc issue peak): runs `iters` dependent-free XOR+POPC+ADD rounds on every SM and
 * returns giga-(32-bit popc)/s in *gpopc_per_s. */
CVB_API int cvb_microbench_popc(cvb_ctx* ctx, int iters, double* gpopc_per_s);

#ifdef __cplusplus
}
#endif
#endif /* COVINS_B200_H_ */

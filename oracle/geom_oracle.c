/* oracle/geom_oracle.c — CPU restatement of the guided search and of the RANSAC hypothesis scoring.
 * TEST INFRASTRUCTURE ONLY (tests/, __graft_entry__.smoke(), bench.py cpu legs); never linked into covins_b200/.
 *
 * PARITY UNPINNED beyond the in-tree sources: the reference has no tests for these stages and its callers need
 * Eigen/OpenCV/opengv (absent).  Restated line by line from
 *   FeatureMatcher::SearchBySE3            covins_backend/src/covins_backend/feature_matcher_be.cpp:293-498
 *   KeyframeBase::GetFeaturesInArea        covins_backend/src/covins_base/keyframe_base.cpp:262-318 (grid branch)
 *   KeyframeBase::IsInImage                keyframe_base.cpp:414-416
 *   LandmarkBase::PredictScale             covins_backend/src/covins_base/landmark_base.cpp:120-133
 *   FeatureMatcher::DescriptorDistanceHamming   feature_matcher_be.cpp:49-64
 *   FrameAbsolutePoseSacProblem::getSelectedDistancesToModel   include/covins/matcher/opengv/sac_problems/FrameAbsolutePoseSacProblem.h:95-126
 *   FrameRelativePoseSacProblem::getSelectedDistancesToModel   include/covins/matcher/opengv/sac_problems/frame-relative-pose-sac-problem.hpp:69-104
 *   opengv::triangulation::triangulate2    [A] opengv (not in the tree, dependencies.rosinstall:43-45): the closed-form
 *                                          two-ray mid-point: lambda = A^-1 b, X = (lambda0 f1 + t12 + lambda1 R12 f2) / 2
 * Arithmetic is plain IEEE double/float without contraction (compile with -ffp-contract=off): the CUDA path uses explicit
 * non-fused operations, so results are compared bit for bit.
 */
#include <float.h>
#include <limits.h>
#include <math.h>
#include <stdint.h>
#include <string.h>

#define API __attribute__((visibility("default")))
#define GRID_COLS 64
#define GRID_ROWS 48

typedef struct ora_kf_view {
  int32_t n;
  const float* kp; const float* octave; const uint8_t* desc; const uint8_t* lm_valid; const double* lm_pos;
  const double* lm_maxdist; const uint8_t* lm_desc; const int32_t* grid_ptr; const int32_t* grid_idx;
  double grid_w_inv, grid_h_inv; double K[9]; double Tcw[16]; double img[4];
} ora_kf_view;
typedef struct ora_search_params { double th; int32_t desc_th_low; int32_t num_octaves; double scale_factor; } ora_search_params;

static int ham256(const uint8_t* a, const uint8_t* b) {
  uint32_t x[8], y[8];
  memcpy(x, a, 32); memcpy(y, b, 32);
  int d = 0;
  for (int i = 0; i < 8; i++) d += __builtin_popcount(x[i] ^ y[i]);
  return d;
}
static void rt_apply(const double* T, const double* p, double* o) {   /* T.block<3,3>(0,0)*p + T.block<3,1>(0,3) */
  for (int r = 0; r < 3; r++) o[r] = (T[4 * r] * p[0] + T[4 * r + 1] * p[1] + T[4 * r + 2] * p[2]) + T[4 * r + 3];
}
static int in_image(const ora_kf_view* k, double x, double y) { return x >= k->img[0] && x < k->img[1] && y >= k->img[2] && y < k->img[3]; }
static int predict_scale(double maxdist, double dist3d, const ora_search_params* prm) {
  const float cur = (float)dist3d;                          /* const float& currentDist */
  const double ratio = maxdist / (double)cur;
  int n = (int)ceil(log(ratio) / log(prm->scale_factor));
  if (n < 0) n = 0; else if (n >= prm->num_octaves) n = prm->num_octaves - 1;
  return n;
}
/* best candidate of `src` landmark i projected with (Ta then Tb) into `dst`, candidates from dst's grid */
static int search_one(const ora_kf_view* src, int i, const double* Tcw_src, const double* Tab, const double* Kdst, const ora_kf_view* dst,
                      const ora_kf_view* img_kf, const ora_search_params* prm, int float_best, int* best_dist_out) {
  double pc_src[3], pc[3];
  rt_apply(Tcw_src, src->lm_pos + 3 * (size_t)i, pc_src);
  rt_apply(Tab, pc_src, pc);
  if (pc[2] < 0.0) return -1;
  double proj[3];
  for (int r = 0; r < 3; r++) proj[r] = Kdst[3 * r] * pc[0] + Kdst[3 * r + 1] * pc[1] + Kdst[3 * r + 2] * pc[2];
  const double u = proj[0] / proj[2], v = proj[1] / proj[2];
  if (!in_image(img_kf, u, v)) return -1;
  const double dist3d = sqrt((pc[0] * pc[0] + pc[1] * pc[1]) + pc[2] * pc[2]);
  const int level = predict_scale(src->lm_maxdist[i], dist3d, prm);
  const double radius = prm->th * pow(2.0, (double)level);
  /* GetFeaturesInArea(target = KeypointType(u, v) — floats —, radius) */
  const float tx = (float)u, ty = (float)v;
  int min_cx = (int)floor(((double)tx - radius) * dst->grid_w_inv); if (min_cx < 0) min_cx = 0;
  if (min_cx >= GRID_COLS) return -1;
  int max_cx = (int)ceil(((double)tx + radius) * dst->grid_w_inv); if (max_cx > GRID_COLS - 1) max_cx = GRID_COLS - 1;
  if (max_cx < 0) return -1;
  int min_cy = (int)floor(((double)ty - radius) * dst->grid_h_inv); if (min_cy < 0) min_cy = 0;
  if (min_cy >= GRID_ROWS) return -1;
  int max_cy = (int)ceil(((double)ty + radius) * dst->grid_h_inv); if (max_cy > GRID_ROWS - 1) max_cy = GRID_ROWS - 1;
  if (max_cy < 0) return -1;
  double best_f = (double)FLT_MAX;   /* float bestDist = FLT_MAX (direction 1) */
  int best_i = INT_MAX;              /* int bestDist = INT_MAX   (direction 2) */
  int best = -1;
  for (int ix = min_cx; ix <= max_cx; ix++)
    for (int iy = min_cy; iy <= max_cy; iy++) {
      const int c = ix * GRID_ROWS + iy;
      for (int q = dst->grid_ptr[c]; q < dst->grid_ptr[c + 1]; q++) {
        const int idx = dst->grid_idx[q];
        const float dx = dst->kp[2 * (size_t)idx] - tx, dy = dst->kp[2 * (size_t)idx + 1] - ty;
        const float nrm = sqrtf(dx * dx + dy * dy);
        if (!((double)nrm <= radius)) continue;
        const int lvl = (int)dst->octave[idx];
        if (lvl < level - 1 || lvl > level) continue;
        const int d = ham256(src->lm_desc + 32 * (size_t)i, dst->desc + 32 * (size_t)idx);
        if (float_best) { if ((double)d < best_f) { best_f = (double)d; best = idx; } }
        else { if (d < best_i) { best_i = d; best = idx; } }
      }
    }
  *best_dist_out = float_best ? (best >= 0 ? (int)best_f : INT_MAX) : best_i;
  return best;
}

API void ora_search_by_se3(const ora_kf_view* kf1, const ora_kf_view* kf2, const double* T12, const double* T21, const uint8_t* already1,
                           const uint8_t* already2, const ora_search_params* prm, int32_t* match12, int32_t* n_found, int32_t* match1,
                           int32_t* match2) {
  const int n1 = kf1->n, n2 = kf2->n;
  for (int i = 0; i < n1; i++) {
    match1[i] = -1;
    if (!kf1->lm_valid[i] || already1[i]) continue;
    int bd;
    const int b = search_one(kf1, i, kf1->Tcw, T21, kf2->K, kf2, kf2, prm, 1, &bd);
    if (b >= 0 && (float)bd <= (float)prm->desc_th_low) match1[i] = b;        /* bestDist <= desc_matching_th_low_ (:403) */
  }
  for (int i = 0; i < n2; i++) {
    match2[i] = -1;
    if (!kf2->lm_valid[i] || already2[i]) continue;
    int bd;
    const int b = search_one(kf2, i, kf2->Tcw, T12, kf1->K, kf1, kf2 /* pKF2->IsInImage, :433 */, prm, 0, &bd);
    if (b >= 0 && bd < prm->desc_th_low) match2[i] = b;                         /* bestDist < desc_matching_th_low_ (:479) */
  }
  int found = 0;
  for (int i = 0; i < n1; i++) {
    match12[i] = -1;
    const int idx2 = match1[i];
    if (idx2 >= 0) {
      const int idx1 = i < n2 ? match2[i] : -1;                                 /* match2[i], not match2[idx2] (:489) */
      if (idx1 == i) { match12[i] = idx2; found++; }
    }
  }
  *n_found = found;
}

/* ---------------------------------------------------------------------------------------------- V1 scoring */
API void ora_score_absolute_pose(const double* model, int n_hyp, const double* pts, const double* f, const double* sigma, int n,
                                 const double* cam_off, const double* cam_rot, double threshold, double* scores, uint8_t* inlier,
                                 int32_t* n_inliers) {
  for (int h = 0; h < n_hyp; h++) {
    const double* M = model + 12 * (size_t)h;     /* 3x4 [R|t] row-major */
    /* inverseSolution = [R^T | -R^T t] */
    double Ri[9], ti[3];
    for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) Ri[3 * r + c] = M[4 * c + r];
    for (int r = 0; r < 3; r++) ti[r] = -((Ri[3 * r] * M[3] + Ri[3 * r + 1] * M[7]) + Ri[3 * r + 2] * M[11]);
    int cnt = 0;
    for (int i = 0; i < n; i++) {
      const double* p = pts + 3 * (size_t)i;
      double b[3], q[3], e2 = 0.0;
      for (int r = 0; r < 3; r++) b[r] = ((Ri[3 * r] * p[0] + Ri[3 * r + 1] * p[1]) + Ri[3 * r + 2] * p[2]) + ti[r];   /* inverseSolution * p_hom */
      for (int r = 0; r < 3; r++) b[r] = b[r] - cam_off[r];
      for (int r = 0; r < 3; r++) q[r] = (cam_rot[r] * b[0] + cam_rot[3 + r] * b[1]) + cam_rot[6 + r] * b[2];          /* Rc^T (.) */
      const double nrm = sqrt((q[0] * q[0] + q[1] * q[1]) + q[2] * q[2]);
      for (int r = 0; r < 3; r++) { const double e = q[r] / nrm - f[3 * (size_t)i + r]; e2 = r == 0 ? e * e : e2 + e * e; }
      const double s = e2 / sigma[i];
      if (scores) scores[(size_t)h * n + i] = s;
      const int in = s < threshold;
      if (inlier) inlier[(size_t)h * n + i] = (uint8_t)in;
      cnt += in;
    }
    n_inliers[h] = cnt;
  }
}

API void ora_score_relative_pose(const double* model, int n_hyp, const double* f1, const double* f2, const double* sigma1,
                                 const double* sigma2, int n, double threshold, double* scores, uint8_t* inlier, int32_t* n_inliers) {
  for (int h = 0; h < n_hyp; h++) {
    const double* M = model + 12 * (size_t)h;
    const double t[3] = {M[3], M[7], M[11]};
    double Ri[9], ti[3];
    for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) Ri[3 * r + c] = M[4 * c + r];
    for (int r = 0; r < 3; r++) ti[r] = -((Ri[3 * r] * t[0] + Ri[3 * r + 1] * t[1]) + Ri[3 * r + 2] * t[2]);
    int cnt = 0;
    for (int i = 0; i < n; i++) {
      const double* a = f1 + 3 * (size_t)i;
      const double* bb = f2 + 3 * (size_t)i;
      double u[3];   /* f2_unrotated = R12 f2 */
      for (int r = 0; r < 3; r++) u[r] = (M[4 * r] * bb[0] + M[4 * r + 1] * bb[1]) + M[4 * r + 2] * bb[2];
      const double b0 = (t[0] * a[0] + t[1] * a[1]) + t[2] * a[2], b1 = (t[0] * u[0] + t[1] * u[1]) + t[2] * u[2];
      const double A00 = (a[0] * a[0] + a[1] * a[1]) + a[2] * a[2], A10 = (a[0] * u[0] + a[1] * u[1]) + a[2] * u[2];
      const double A01 = -A10, A11 = -((u[0] * u[0] + u[1] * u[1]) + u[2] * u[2]);
      const double det = A00 * A11 - A01 * A10;
      /* lambda = A^-1 b (2x2 inverse by the adjugate) */
      const double l0 = (A11 * b0 - A01 * b1) / det, l1 = (A00 * b1 - A10 * b0) / det;
      double X[3], r2[3];
      for (int r = 0; r < 3; r++) X[r] = (l0 * a[r] + (t[r] + l1 * u[r])) / 2.0;
      for (int r = 0; r < 3; r++) r2[r] = ((Ri[3 * r] * X[0] + Ri[3 * r + 1] * X[1]) + Ri[3 * r + 2] * X[2]) + ti[r];
      const double n1 = sqrt((X[0] * X[0] + X[1] * X[1]) + X[2] * X[2]), n2 = sqrt((r2[0] * r2[0] + r2[1] * r2[1]) + r2[2] * r2[2]);
      double e1 = 0.0, e2 = 0.0;
      for (int r = 0; r < 3; r++) {
        const double d1 = X[r] / n1 - a[r], d2 = r2[r] / n2 - bb[r];
        e1 = r == 0 ? d1 * d1 : e1 + d1 * d1;
        e2 = r == 0 ? d2 * d2 : e2 + d2 * d2;
      }
      const double s = e1 * 0.5 / sigma1[i] + e2 * 0.5 / sigma2[i];
      if (scores) scores[(size_t)h * n + i] = s;
      const int in = s < threshold;
      if (inlier) inlier[(size_t)h * n + i] = (uint8_t)in;
      cnt += in;
    }
    n_inliers[h] = cnt;
  }
}

/* ---------------------------------------------------------------------------------------------- SearchByProjection
 * FeatureMatcher::SearchByProjection (feature_matcher_be.cpp:168-291): candidate landmarks are projected into the keyframe
 * with its full camera model (camera_->project3), gated (image, distance invariance, viewing angle), matched to the best
 * free keypoint in a radius, and the outcome is applied IN ORDER: a keypoint taken by an earlier landmark (vpMatched) is
 * skipped by the later ones, RemapLandmark (keyframe_be.cpp:484-495) moves an already-observed landmark to a better
 * keypoint and erases the observation of the landmark it displaces. */
typedef struct ora_proj_landmarks {
  int32_t m;
  const uint8_t* valid; const double* pos; const double* normal; const double* min_dist; const double* max_dist;
  const double* max_distance; const uint8_t* desc; const int32_t* feat_idx;
} ora_proj_landmarks;

static int project3(const double* intr, const double* dist, int cam_model, int dist_model, double xi, const double* pc, double* u, double* v) {
  double den = pc[2];
  if (cam_model == 1) den = pc[2] + xi * sqrt(pc[0] * pc[0] + pc[1] * pc[1] + pc[2] * pc[2]);
  if (!(den > 1e-10)) return 0;
  const double x = pc[0] / den, y = pc[1] / den, r2 = x * x + y * y;
  double xd, yd;
  if (dist_model == 0) {
    const double rad = 1.0 + dist[0] * r2 + dist[1] * r2 * r2;
    xd = x * rad + 2.0 * dist[2] * x * y + dist[3] * (r2 + 2.0 * x * x);
    yd = y * rad + dist[2] * (r2 + 2.0 * y * y) + 2.0 * dist[3] * x * y;
  } else if (dist_model == 1) {
    const double r = sqrt(r2);
    double s = 1.0;
    if (r >= 1e-8) { const double th = atan(r), t2 = th * th; s = th * (1.0 + dist[0] * t2 + dist[1] * t2 * t2 + dist[2] * t2 * t2 * t2 + dist[3] * t2 * t2 * t2 * t2) / r; }
    xd = s * x; yd = s * y;
  } else {
    const double w = dist[0];
    double s = 1.0;
    if (w * w >= 1e-5) { const double c = 2.0 * tan(0.5 * w); s = r2 < 1e-5 ? c / w : atan(c * sqrt(r2)) / (w * sqrt(r2)); }
    xd = s * x; yd = s * y;
  }
  *u = intr[0] * xd + intr[2]; *v = intr[1] * yd + intr[3];
  return 1;
}

API void ora_search_by_projection(const ora_kf_view* kf, const int32_t* kf_lm_cand_in, const double* Tcw, const double* intr, const double* dist,
                                  int cam_model, int dist_model, double xi, const ora_proj_landmarks* L, const uint8_t* matched_in,
                                  const ora_search_params* prm, int32_t* action, int32_t* best_idx, int32_t* n_matches) {
  const int n = kf->n, m = L->m;
  uint8_t* matched = (uint8_t*)__builtin_alloca(n > 0 ? n : 1);
  uint8_t* has_lm = (uint8_t*)__builtin_alloca(n > 0 ? n : 1);
  int32_t* lm_cand = (int32_t*)__builtin_alloca(sizeof(int32_t) * (n > 0 ? n : 1));
  int32_t* feat = (int32_t*)__builtin_alloca(sizeof(int32_t) * (m > 0 ? m : 1));
  for (int i = 0; i < n; i++) { matched[i] = matched_in[i]; has_lm[i] = kf->lm_valid[i]; lm_cand[i] = kf_lm_cand_in[i]; }
  for (int i = 0; i < m; i++) feat[i] = L->feat_idx[i];
  /* Ow = -Rcw^T tcw */
  double Ow[3];
  for (int r = 0; r < 3; r++) Ow[r] = -((Tcw[r] * Tcw[3] + Tcw[4 + r] * Tcw[7]) + Tcw[8 + r] * Tcw[11]);
  int nm = 0;
  for (int i = 0; i < m; i++) {
    action[i] = 0; best_idx[i] = -1;
    if (!L->valid[i]) continue;
    const double* pw = L->pos + 3 * (size_t)i;
    double pc[3];
    rt_apply(Tcw, pw, pc);
    if (pc[2] < 0.0) continue;
    double u, v;
    if (!project3(intr, dist, cam_model, dist_model, xi, pc, &u, &v)) continue;
    if (!in_image(kf, u, v)) continue;
    const double PO[3] = {pw[0] - Ow[0], pw[1] - Ow[1], pw[2] - Ow[2]};
    const double d3 = sqrt((PO[0] * PO[0] + PO[1] * PO[1]) + PO[2] * PO[2]);
    if (d3 < L->min_dist[i] || d3 > L->max_dist[i]) continue;
    const double* Pn = L->normal + 3 * (size_t)i;
    if ((PO[0] * Pn[0] + PO[1] * Pn[1]) + PO[2] * Pn[2] < 0.5 * d3) continue;
    const int level = predict_scale(L->max_distance[i], d3, prm);
    const double radius = prm->th * pow(prm->scale_factor, (double)level);
    const float tx = (float)u, ty = (float)v;
    int min_cx = (int)floor(((double)tx - radius) * kf->grid_w_inv); if (min_cx < 0) min_cx = 0;
    if (min_cx >= GRID_COLS) continue;
    int max_cx = (int)ceil(((double)tx + radius) * kf->grid_w_inv); if (max_cx > GRID_COLS - 1) max_cx = GRID_COLS - 1;
    if (max_cx < 0) continue;
    int min_cy = (int)floor(((double)ty - radius) * kf->grid_h_inv); if (min_cy < 0) min_cy = 0;
    if (min_cy >= GRID_ROWS) continue;
    int max_cy = (int)ceil(((double)ty + radius) * kf->grid_h_inv); if (max_cy > GRID_ROWS - 1) max_cy = GRID_ROWS - 1;
    if (max_cy < 0) continue;
    int bd = 256, best = -1;
    const uint8_t* dl = L->desc + 32 * (size_t)i;
    for (int ix = min_cx; ix <= max_cx; ix++)
      for (int iy = min_cy; iy <= max_cy; iy++) {
        const int c = ix * GRID_ROWS + iy;
        for (int q = kf->grid_ptr[c]; q < kf->grid_ptr[c + 1]; q++) {
          const int idx = kf->grid_idx[q];
          const float dx = kf->kp[2 * (size_t)idx] - tx, dy = kf->kp[2 * (size_t)idx + 1] - ty;
          if (!((double)sqrtf(dx * dx + dy * dy) <= radius)) continue;
          if (matched[idx]) continue;                                         /* if (vpMatched[idx]) continue (:239) */
          const int lvl = (int)kf->octave[idx];
          if (lvl < level - 1 || lvl > level) continue;
          const int d = ham256(dl, kf->desc + 32 * (size_t)idx);
          if (d < bd) { bd = d; best = idx; }
        }
      }
    if (best < 0 || bd > prm->desc_th_low) continue;                          /* bestDist <= desc_matching_th_low_ (:258) */
    best_idx[i] = best;
    const int existing = feat[i];
    if (existing != -1) {                                                     /* already observed (:260-282) */
      int keep = 0;
      if (ham256(dl, kf->desc + 32 * (size_t)existing) < bd) keep = 1;
      if (has_lm[best] && ham256(dl, kf->desc + 32 * (size_t)best) < bd) keep = 1;
      if (keep) { action[i] = 3; continue; }
      /* RemapLandmark(pMP, existing, best) (keyframe_be.cpp:484-495) */
      const int displaced = lm_cand[best];
      const int had = has_lm[best];
      has_lm[existing] = 0; lm_cand[existing] = -1;
      has_lm[best] = 1; lm_cand[best] = i; feat[i] = best;
      if (had && displaced >= 0) feat[displaced] = -1;                        /* lm_new->EraseObservation(kf) is the LAST statement (:494) */
      action[i] = 2;
    } else {
      matched[best] = 1;                                                      /* vpMatched[bestIdx] = pMP (:285) */
      action[i] = 1; nm++;
    }
  }
  *n_matches = nm;
}

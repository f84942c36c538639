/* CPU test double of the C-ABI for tests/test_shim_host.py: the shim's SearchByProjection wrapper is host logic (flatten the
 * containers, replay the decisions); here the library call behind it is answered by the C oracle (oracle/geom_oracle.c), so the
 * wrapper can be exercised without a GPU.  Test infrastructure only — the product library has no CPU path. */
#include <stdint.h>
#include <stddef.h>
#include "../../include/covins_b200.h"

typedef struct ora_kf_view ora_kf_view;
typedef struct ora_proj_landmarks ora_proj_landmarks;
typedef struct ora_search_params ora_search_params;
/* same field order and types as cvb_kf_view / cvb_proj_landmarks / cvb_search_params (oracle/geom_oracle.c:28-34,205-209) */
void ora_search_by_projection(const ora_kf_view* kf, const int32_t* kf_lm_cand_in, const double* Tcw, const double* intr, const double* dist,
                              int cam_model, int dist_model, double xi, const ora_proj_landmarks* lms, const uint8_t* matched_in,
                              const ora_search_params* prm, int32_t* action, int32_t* best_idx, int32_t* n_matches);

static int dummy_ctx;
int cvb_ctx_create(int device, cvb_ctx** out) { (void)device; *out = (cvb_ctx*)&dummy_ctx; return CVB_OK; }
int cvb_ctx_destroy(cvb_ctx* ctx) { (void)ctx; return CVB_OK; }
const char* cvb_last_error(const cvb_ctx* ctx) { (void)ctx; return ""; }
int cvb_search_by_projection(cvb_ctx* ctx, const cvb_kf_view* kf, const int32_t* kf_lm_cand, const double* Tcw, const double* intr,
                             const double* dist, int cam_model, int dist_model, double xi, const cvb_proj_landmarks* lms,
                             const uint8_t* matched, const cvb_search_params* prm, int32_t* action, int32_t* best_idx,
                             int32_t* n_matches) {
  (void)ctx;
  ora_search_by_projection((const ora_kf_view*)kf, kf_lm_cand, Tcw, intr, dist, cam_model, dist_model, xi, (const ora_proj_landmarks*)lms, matched,
                           (const ora_search_params*)prm, action, best_idx, n_matches);
  return CVB_OK;
}

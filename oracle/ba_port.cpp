// oracle/ba_port.cpp — C++17 / OpenMP CPU port of the reference's optimisation hot path (global BA / PGO).
//
// TEST INFRASTRUCTURE + the timed CPU baseline (bench.py `cpu_baseline` / `--impl reference`); never linked into
// covins_b200/.  SURVEY.md §8(d) "CPU baseline timing": the reference binary (Ceres + CHOLMOD + robopt_open) cannot be
// built here, so its CPU path is restated as a compiled, threaded program with the structure of Ceres' SPARSE_SCHUR:
//   cost functions with ANALYTIC Jacobians (robopt_open style), CauchyLoss + corrector, Jacobi scaling,
//   landmark elimination (Schur complement) into a block-sparse reduced camera system,
//   a supernodal-style sparse Cholesky (128-wide dense tiles with symbolic fill, BLAS-3 kernels from the OpenBLAS that
//   ships with scipy — the role CHOLMOD + BLAS play under Ceres), back-substitution,
//   Ceres 1.x TrustRegionMinimizer + traditional DoglegStrategy.
// It follows oracle/ba_oracle.py (the autograd restatement; assumptions [A] listed there and in SURVEY Appendix A)
// step for step and is checked against it to ~1e-9 (tests/test_ba_port.py); every citation of the reference's call sites
// is in ba_oracle.py's header: optimization_be.cpp:56-618 (GBA), :833-1086 (PGO).
//
// Ordering of the reduced camera system: speed-bias blocks first, IMU chain by IMU chain, then pose blocks chain by
// chain (each chain starts on a tile boundary).  Along a chain the speed-bias part is block-banded, so its elimination
// is cheap and the fill stays inside the pose part — the same observation an AMD ordering makes for CHOLMOD.
#include <dlfcn.h>
#include <omp.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <functional>
#include <numeric>
#include <vector>

#define BAP_API extern "C" __attribute__((visibility("default")))

extern "C" {
struct bap_problem {   // same flat format as include/covins_b200.h:cvb_ba_problem
  int32_t K, L, n_obs, n_imu, n_edge, n_cam;
  const double* pose; const double* speedbias; const uint8_t* pose_const; const int32_t* cam_of_kf;
  const double* extr; const double* intr; const double* dist;
  const double* lm; const int32_t* lm_obs_ptr; const int32_t* obs_kf; const float* obs_uv; const double* obs_sigma;
  const uint8_t* obs_skip;
  const int32_t* imu_i; const int32_t* imu_j; const int32_t* imu_ptr; const double* imu_dt; const double* imu_acc;
  const double* imu_gyr; const double* imu_acc0; const double* imu_gyr0; const double* imu_noise;
  const int32_t* edge_i; const int32_t* edge_j; const double* edge_q; const double* edge_t; const double* edge_sqrt_info;
  const uint8_t* edge_robust;
};
struct bap_options {
  int32_t max_iterations, visual_only;
  double cauchy_reproj, cauchy_edge;
  int32_t threads;   // <= 0: omp default
};
struct bap_result {
  double* pose; double* speedbias; double* lm;
  double* cost_history; uint8_t* step_status; int32_t cost_history_cap, n_cost_history;
  int32_t iterations, termination;
  double initial_cost, final_cost;
  double phase_s[6];   // linearise, blocks+Schur, factor, solve+backsub, dogleg/step/candidate cost, set-up
  double factor_flops; // executed tile-GEMM flops per factorisation
};
}

namespace {

// ---------------------------------------------------------------------------------------------- BLAS (dlopen'ed)
typedef void (*dgemm_t)(int, int, int, int, int, int, double, const double*, int, const double*, int, double, double*, int);
typedef void (*dsyrk_t)(int, int, int, int, int, double, const double*, int, double, double*, int);
typedef void (*dtrsm_t)(int, int, int, int, int, int, int, double, const double*, int, double*, int);
typedef int (*dpotrf_t)(int, char, int, double*, int);
typedef void (*setthr_t)(int);
dgemm_t p_dgemm = nullptr;
dsyrk_t p_dsyrk = nullptr;
dtrsm_t p_dtrsm = nullptr;
dpotrf_t p_dpotrf = nullptr;
setthr_t p_setthr = nullptr;
enum { RowMajor = 101, NoTrans = 111, Trans = 112, Upper = 121, Lower = 122, NonUnit = 131, Left = 141, Right = 142 };

constexpr int T = 128;

// C(T x T) -= A(T x T) B(T x T)^T, row-major tiles
void tile_gemm_nt(const double* A, const double* B, double* C) {
  if (p_dgemm) { p_dgemm(RowMajor, NoTrans, Trans, T, T, T, -1.0, A, T, B, T, 1.0, C, T); return; }
  for (int i = 0; i < T; i++)
    for (int j = 0; j < T; j++) {
      double s = 0;
      for (int k = 0; k < T; k++) s += A[i * T + k] * B[j * T + k];
      C[i * T + j] -= s;
    }
}
void tile_syrk(const double* A, double* C) {   // lower(C) -= A A^T
  if (p_dsyrk) { p_dsyrk(RowMajor, Lower, NoTrans, T, T, -1.0, A, T, 1.0, C, T); return; }
  for (int i = 0; i < T; i++)
    for (int j = 0; j <= i; j++) {
      double s = 0;
      for (int k = 0; k < T; k++) s += A[i * T + k] * A[j * T + k];
      C[i * T + j] -= s;
    }
}
bool tile_potrf(double* A) {   // lower Cholesky in place (upper part left untouched)
  if (p_dpotrf) return p_dpotrf(RowMajor, 'L', T, A, T) == 0;
  for (int j = 0; j < T; j++) {
    double d = A[j * T + j];
    for (int k = 0; k < j; k++) d -= A[j * T + k] * A[j * T + k];
    if (!(d > 0)) return false;
    d = std::sqrt(d);
    A[j * T + j] = d;
    for (int i = j + 1; i < T; i++) {
      double s = A[i * T + j];
      for (int k = 0; k < j; k++) s -= A[i * T + k] * A[j * T + k];
      A[i * T + j] = s / d;
    }
  }
  return true;
}
void tile_trsm(const double* Lkk, double* A) {   // A <- A Lkk^-T
  if (p_dtrsm) { p_dtrsm(RowMajor, Right, Lower, Trans, NonUnit, T, T, 1.0, Lkk, T, A, T); return; }
  for (int i = 0; i < T; i++)
    for (int j = 0; j < T; j++) {
      double s = A[i * T + j];
      for (int k = 0; k < j; k++) s -= A[i * T + k] * Lkk[j * T + k];
      A[i * T + j] = s / Lkk[j * T + j];
    }
}

// ---------------------------------------------------------------------------------------------- small math
struct V3 { double x, y, z; };
inline V3 operator+(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline V3 operator*(double s, V3 a) { return {s * a.x, s * a.y, s * a.z}; }
inline double dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
struct M3 { double m[9]; };
inline M3 eye() { return {{1, 0, 0, 0, 1, 0, 0, 0, 1}}; }
inline M3 skew(V3 v) { return {{0, -v.z, v.y, v.z, 0, -v.x, -v.y, v.x, 0}}; }
inline M3 tr(const M3& A) { return {{A.m[0], A.m[3], A.m[6], A.m[1], A.m[4], A.m[7], A.m[2], A.m[5], A.m[8]}}; }
inline M3 mm(const M3& A, const M3& B) {
  M3 C;
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) C.m[3 * i + j] = A.m[3 * i] * B.m[j] + A.m[3 * i + 1] * B.m[3 + j] + A.m[3 * i + 2] * B.m[6 + j];
  return C;
}
inline V3 mv(const M3& A, V3 v) {
  return {A.m[0] * v.x + A.m[1] * v.y + A.m[2] * v.z, A.m[3] * v.x + A.m[4] * v.y + A.m[5] * v.z, A.m[6] * v.x + A.m[7] * v.y + A.m[8] * v.z};
}
inline V3 mtv(const M3& A, V3 v) {
  return {A.m[0] * v.x + A.m[3] * v.y + A.m[6] * v.z, A.m[1] * v.x + A.m[4] * v.y + A.m[7] * v.z, A.m[2] * v.x + A.m[5] * v.y + A.m[8] * v.z};
}
struct Q { double x, y, z, w; };
inline Q qm(Q a, Q b) {
  return {a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y, a.w * b.y - a.x * b.z + a.y * b.w + a.z * b.x,
          a.w * b.z + a.x * b.y - a.y * b.x + a.z * b.w, a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z};
}
inline Q qc(Q q) { return {-q.x, -q.y, -q.z, q.w}; }
inline Q qn(Q q) { const double n = 1.0 / std::sqrt(q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w); return {q.x * n, q.y * n, q.z * n, q.w * n}; }
inline M3 q2R(Q q) {
  const double x = q.x, y = q.y, z = q.z, w = q.w;
  return {{1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w), 2 * (x * y + z * w), 1 - 2 * (x * x + z * z),
           2 * (y * z - x * w), 2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)}};
}
inline Q qexp(V3 p) {
  const double th2 = dot(p, p);
  double k, w;
  if (th2 < 1e-12) { k = 0.5 - th2 / 48.0; w = 1.0 - th2 / 8.0; }
  else { const double th = std::sqrt(th2); k = std::sin(0.5 * th) / th; w = std::cos(0.5 * th); }
  return {k * p.x, k * p.y, k * p.z, w};
}
inline M3 so3_Jr(V3 p) {   // right Jacobian of SO(3)
  const double th2 = dot(p, p);
  const M3 Kx = skew(p), K2 = mm(Kx, Kx);
  double a, b;
  if (th2 < 1e-10) { a = 0.5 - th2 / 24.0; b = 1.0 / 6.0 - th2 / 120.0; }
  else { const double th = std::sqrt(th2); a = (1.0 - std::cos(th)) / th2; b = (th - std::sin(th)) / (th2 * th); }
  M3 J = eye();
  for (int i = 0; i < 9; i++) J.m[i] += -a * Kx.m[i] + b * K2.m[i];
  return J;
}
inline M3 quat_rjac(Q E) {   // d(2 vec(E Exp(phi)))/dphi at 0
  M3 J = skew({E.x, E.y, E.z});
  J.m[0] += E.w; J.m[4] += E.w; J.m[8] += E.w;
  return J;
}
inline void cauchy(double s, double a2, double* scale, double* cost) {
  if (a2 <= 0) { *scale = 1.0; *cost = 0.5 * s; }
  else { *scale = std::sqrt(1.0 / (1.0 + s / a2)); *cost = 0.5 * a2 * std::log1p(s / a2); }
}
inline void pose_plus(const double* p, const double* d, double* o) {
  const Q q = qn(qm(qexp({d[0], d[1], d[2]}), {p[0], p[1], p[2], p[3]}));
  o[0] = q.x; o[1] = q.y; o[2] = q.z; o[3] = q.w;
  o[4] = p[4] + d[3]; o[5] = p[5] + d[4]; o[6] = p[6] + d[5];
}

// reprojection residual (2) with Jacobians w.r.t. [dtheta, dp] (2x6) and the landmark (2x3)
inline void f_reproj(const double* pose, const double* extr, const double* intr, const double* dist, const double* lm, double u,
                     double v, double sigma, double r[2], double* Jp, double* Jl) {
  const M3 Rws = q2R({pose[0], pose[1], pose[2], pose[3]}), Rsc = q2R({extr[0], extr[1], extr[2], extr[3]});
  const V3 d{lm[0] - pose[4], lm[1] - pose[5], lm[2] - pose[6]};
  const V3 ps = mtv(Rws, d);
  const V3 pc = mtv(Rsc, ps - V3{extr[4], extr[5], extr[6]});
  if (!(pc.z > 1e-10)) {
    r[0] = r[1] = 0;
    if (Jp) { std::fill(Jp, Jp + 12, 0.0); std::fill(Jl, Jl + 6, 0.0); }
    return;
  }
  const double iz = 1.0 / pc.z, x = pc.x * iz, y = pc.y * iz;
  const double k1 = dist[0], k2 = dist[1], p1 = dist[2], p2 = dist[3];
  const double r2 = x * x + y * y, rad = 1 + k1 * r2 + k2 * r2 * r2;
  const double xd = x * rad + 2 * p1 * x * y + p2 * (r2 + 2 * x * x), yd = y * rad + p1 * (r2 + 2 * y * y) + 2 * p2 * x * y;
  const double is = 1.0 / sigma;
  r[0] = (intr[0] * xd + intr[2] - u) * is;
  r[1] = (intr[1] * yd + intr[3] - v) * is;
  if (!Jp) return;
  const double c = k1 + 2 * k2 * r2;
  const double dxx = rad + 2 * x * x * c + 2 * p1 * y + 6 * p2 * x, dxy = 2 * x * y * c + 2 * p1 * x + 2 * p2 * y;
  const double dyx = dxy, dyy = rad + 2 * y * y * c + 6 * p1 * y + 2 * p2 * x;
  const double fx = intr[0] * is, fy = intr[1] * is;
  const double A[6] = {fx * dxx * iz, fx * dxy * iz, -fx * (dxx * x + dxy * y) * iz, fy * dyx * iz, fy * dyy * iz, -fy * (dyx * x + dyy * y) * iz};
  // d pc / d lm = Rsc^T Rws^T
  const M3 G = mm(tr(Rsc), tr(Rws));
  for (int a = 0; a < 2; a++)
    for (int b = 0; b < 3; b++) Jl[3 * a + b] = A[3 * a] * G.m[b] + A[3 * a + 1] * G.m[3 + b] + A[3 * a + 2] * G.m[6 + b];
  for (int a = 0; a < 2; a++) {
    const double* j = Jl + 3 * a;
    double* o = Jp + 6 * a;
    o[0] = j[1] * d.z - j[2] * d.y; o[1] = j[2] * d.x - j[0] * d.z; o[2] = j[0] * d.y - j[1] * d.x;
    o[3] = -j[0]; o[4] = -j[1]; o[5] = -j[2];
  }
}

// between factor: e (6) = S [2 vec(qm^-1 q1^-1 q2); R1^T (t2 - t1) - tm], J (6 x 12) = [d/dx1 | d/dx2]
inline void f_between(const double* p1, const double* p2, const double* qmeas, const double* tmeas, const double* S, double e[6], double* J) {
  const Q q1{p1[0], p1[1], p1[2], p1[3]}, q2{p2[0], p2[1], p2[2], p2[3]};
  const Q E = qm(qc({qmeas[0], qmeas[1], qmeas[2], qmeas[3]}), qm(qc(q1), q2));
  const M3 R1 = q2R(q1), R2 = q2R(q2);
  const V3 dt{p2[4] - p1[4], p2[5] - p1[5], p2[6] - p1[6]};
  const V3 et = mtv(R1, dt) - V3{tmeas[0], tmeas[1], tmeas[2]};
  const double raw[6] = {2 * E.x, 2 * E.y, 2 * E.z, et.x, et.y, et.z};
  for (int i = 0; i < 6; i++) { double s = 0; for (int k = 0; k < 6; k++) s += S[6 * i + k] * raw[k]; e[i] = s; }
  if (!J) return;
  const M3 A = mm(quat_rjac(E), tr(R2)), R1t = tr(R1), B = mm(R1t, skew(dt));
  double Jr[72] = {0};
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) {
      Jr[12 * i + j] = -A.m[3 * i + j]; Jr[12 * i + 6 + j] = A.m[3 * i + j];
      Jr[12 * (3 + i) + j] = B.m[3 * i + j]; Jr[12 * (3 + i) + 3 + j] = -R1t.m[3 * i + j]; Jr[12 * (3 + i) + 9 + j] = R1t.m[3 * i + j];
    }
  for (int i = 0; i < 6; i++)
    for (int c = 0; c < 12; c++) { double s = 0; for (int k = 0; k < 6; k++) s += S[6 * i + k] * Jr[12 * k + c]; J[12 * i + c] = s; }
}

struct Pre {   // preintegration of one factor (VINS-Mono style midpoint) [A]
  double Tsum, alpha[3], beta[3], gamma[4], ba[3], bg[3];
  double Jpa[9], Jpg[9], Jqg[9], Jva[9], Jvg[9];
  double W[225];   // sqrt_info = chol(P^-1)^T (upper triangular)
};

void mat_mul(const double* A, const double* B, double* C, int n, int m, int p) {   // C(n x p) = A(n x m) B(m x p)
  for (int i = 0; i < n; i++)
    for (int j = 0; j < p; j++) { double s = 0; for (int k = 0; k < m; k++) s += A[i * m + k] * B[k * p + j]; C[i * p + j] = s; }
}

bool repropagate(const double* dt, const double* acc, const double* gyr, int n, const double* acc0, const double* gyr0, const double* ba_,
                 const double* bg_, const double* noise, Pre& O) {
  const V3 ba{ba_[0], ba_[1], ba_[2]}, bg{bg_[0], bg_[1], bg_[2]};
  const double q_[6] = {noise[0] * noise[0], noise[1] * noise[1], noise[0] * noise[0], noise[1] * noise[1], noise[2] * noise[2], noise[3] * noise[3]};
  std::vector<double> Jm(225, 0.0), P(225, 0.0), F(225), V(270), Tm(225), T2(225);
  for (int i = 0; i < 15; i++) Jm[16 * i] = 1.0;
  V3 dp{0, 0, 0}, dv{0, 0, 0};
  Q dq{0, 0, 0, 1};
  V3 a0{acc0[0], acc0[1], acc0[2]}, g0{gyr0[0], gyr0[1], gyr0[2]};
  double Ts = 0;
  for (int s = 0; s < n; s++) {
    const double h = dt[s];
    const V3 a1{acc[3 * s], acc[3 * s + 1], acc[3 * s + 2]}, g1{gyr[3 * s], gyr[3 * s + 1], gyr[3 * s + 2]};
    const M3 R0 = q2R(dq);
    const V3 ua0 = mv(R0, a0 - ba), ug = 0.5 * (g0 + g1) - bg;
    const Q q1 = qn(qm(dq, {ug.x * h / 2, ug.y * h / 2, ug.z * h / 2, 1.0}));
    const M3 R1 = q2R(q1);
    const V3 ua1 = mv(R1, a1 - ba), ua = 0.5 * (ua0 + ua1);
    const V3 ndp = dp + h * dv + (0.5 * h * h) * ua, ndv = dv + h * ua;
    const M3 Rw = skew(ug), Ra0 = skew(a0 - ba), Ra1 = skew(a1 - ba);
    M3 ImRw = eye();
    for (int i = 0; i < 9; i++) ImRw.m[i] -= Rw.m[i] * h;
    const M3 R0Ra0 = mm(R0, Ra0), R1Ra1 = mm(R1, Ra1), R1Ra1I = mm(R1Ra1, ImRw);
    std::fill(F.begin(), F.end(), 0.0); std::fill(V.begin(), V.end(), 0.0);
    for (int a = 0; a < 3; a++)
      for (int b = 0; b < 3; b++) {
        const int ab = 3 * a + b;
        const double I = a == b ? 1.0 : 0.0;
        F[15 * a + b] = I;
        F[15 * a + 3 + b] = -0.25 * R0Ra0.m[ab] * h * h - 0.25 * R1Ra1I.m[ab] * h * h;
        F[15 * a + 6 + b] = I * h;
        F[15 * a + 9 + b] = -0.25 * (R0.m[ab] + R1.m[ab]) * h * h;
        F[15 * a + 12 + b] = -0.25 * R1Ra1.m[ab] * h * h * (-h);
        F[15 * (3 + a) + 3 + b] = ImRw.m[ab];
        F[15 * (3 + a) + 12 + b] = -I * h;
        F[15 * (6 + a) + 3 + b] = -0.5 * R0Ra0.m[ab] * h - 0.5 * R1Ra1I.m[ab] * h;
        F[15 * (6 + a) + 6 + b] = I;
        F[15 * (6 + a) + 9 + b] = -0.5 * (R0.m[ab] + R1.m[ab]) * h;
        F[15 * (6 + a) + 12 + b] = -0.5 * R1Ra1.m[ab] * h * (-h);
        F[15 * (9 + a) + 9 + b] = I;
        F[15 * (12 + a) + 12 + b] = I;
        V[18 * a + b] = 0.25 * R0.m[ab] * h * h;
        V[18 * a + 3 + b] = 0.25 * (-R1Ra1.m[ab] * h * h) * 0.5 * h;
        V[18 * a + 6 + b] = 0.25 * R1.m[ab] * h * h;
        V[18 * a + 9 + b] = V[18 * a + 3 + b];
        V[18 * (3 + a) + 3 + b] = 0.5 * I * h;
        V[18 * (3 + a) + 9 + b] = 0.5 * I * h;
        V[18 * (6 + a) + b] = 0.5 * R0.m[ab] * h;
        V[18 * (6 + a) + 3 + b] = 0.5 * (-R1Ra1.m[ab] * h) * 0.5 * h;
        V[18 * (6 + a) + 6 + b] = 0.5 * R1.m[ab] * h;
        V[18 * (6 + a) + 9 + b] = V[18 * (6 + a) + 3 + b];
        V[18 * (9 + a) + 12 + b] = I * h;
        V[18 * (12 + a) + 15 + b] = I * h;
      }
    mat_mul(F.data(), Jm.data(), Tm.data(), 15, 15, 15); Jm = Tm;
    mat_mul(F.data(), P.data(), Tm.data(), 15, 15, 15);
    for (int r = 0; r < 15; r++)
      for (int c = 0; c < 15; c++) {
        double s2 = 0;
        for (int m = 0; m < 15; m++) s2 += Tm[15 * r + m] * F[15 * c + m];
        for (int m = 0; m < 18; m++) s2 += V[18 * r + m] * q_[m / 3] * V[18 * c + m];
        T2[15 * r + c] = s2;
      }
    P = T2;
    dp = ndp; dv = ndv; dq = q1; a0 = a1; g0 = g1; Ts += h;
  }
  O.Tsum = Ts;
  O.alpha[0] = dp.x; O.alpha[1] = dp.y; O.alpha[2] = dp.z;
  O.beta[0] = dv.x; O.beta[1] = dv.y; O.beta[2] = dv.z;
  O.gamma[0] = dq.x; O.gamma[1] = dq.y; O.gamma[2] = dq.z; O.gamma[3] = dq.w;
  for (int c = 0; c < 3; c++) { O.ba[c] = ba_[c]; O.bg[c] = bg_[c]; }
  for (int a = 0; a < 3; a++)
    for (int b = 0; b < 3; b++) {
      O.Jpa[3 * a + b] = Jm[15 * a + 9 + b]; O.Jpg[3 * a + b] = Jm[15 * a + 12 + b]; O.Jqg[3 * a + b] = Jm[15 * (3 + a) + 12 + b];
      O.Jva[3 * a + b] = Jm[15 * (6 + a) + 9 + b]; O.Jvg[3 * a + b] = Jm[15 * (6 + a) + 12 + b];
    }
  // sqrt_info = chol(P^-1)^T: P = Lp Lp^T → X = Lp^-1, P^-1 = X^T X, its lower Cholesky factor M, W = M^T
  bool ok = true;
  auto chol = [&](const std::vector<double>& A, std::vector<double>& Lo) {
    std::fill(Lo.begin(), Lo.end(), 0.0);
    for (int c = 0; c < 15; c++)
      for (int r = c; r < 15; r++) {
        double s2 = A[15 * r + c];
        for (int m = 0; m < c; m++) s2 -= Lo[15 * r + m] * Lo[15 * c + m];
        if (r == c) { if (!(s2 > 0)) { ok = false; s2 = 1.0; } Lo[15 * c + c] = std::sqrt(s2); }
        else Lo[15 * r + c] = s2 / Lo[15 * c + c];
      }
  };
  chol(P, F);
  std::fill(Tm.begin(), Tm.end(), 0.0);
  for (int c = 0; c < 15; c++)
    for (int r = c; r < 15; r++) {
      double s2 = r == c ? 1.0 : 0.0;
      for (int m = c; m < r; m++) s2 -= F[15 * r + m] * Tm[15 * m + c];
      Tm[15 * r + c] = s2 / F[15 * r + r];
    }
  for (int r = 0; r < 15; r++)
    for (int c = 0; c < 15; c++) { double s2 = 0; for (int m = std::max(r, c); m < 15; m++) s2 += Tm[15 * m + r] * Tm[15 * m + c]; P[15 * r + c] = s2; }
  chol(P, F);
  for (int r = 0; r < 15; r++)
    for (int c = 0; c < 15; c++) O.W[15 * r + c] = c >= r ? F[15 * c + r] : 0.0;
  return ok;
}

// IMU factor: whitened residual (15), whitened Jacobian (15 x 30: pose_i 6 | sb_i 9 | pose_j 6 | sb_j 9)
void f_imu(const double* pi_, const double* sbi, const double* pj_, const double* sbj, const Pre& P, double g, double rw[15], double* Jw) {
  const Q qi{pi_[0], pi_[1], pi_[2], pi_[3]}, qj{pj_[0], pj_[1], pj_[2], pj_[3]};
  const M3 Ri = q2R(qi), Rj = q2R(qj);
  const V3 ti{pi_[4], pi_[5], pi_[6]}, tj{pj_[4], pj_[5], pj_[6]};
  const V3 vi{sbi[0], sbi[1], sbi[2]}, vj{sbj[0], sbj[1], sbj[2]};
  const V3 dba{sbi[3] - P.ba[0], sbi[4] - P.ba[1], sbi[5] - P.ba[2]}, dbg{sbi[6] - P.bg[0], sbi[7] - P.bg[1], sbi[8] - P.bg[2]};
  M3 Jpa, Jpg, Jqg, Jva, Jvg;
  std::memcpy(Jpa.m, P.Jpa, 72); std::memcpy(Jpg.m, P.Jpg, 72); std::memcpy(Jqg.m, P.Jqg, 72); std::memcpy(Jva.m, P.Jva, 72); std::memcpy(Jvg.m, P.Jvg, 72);
  const double Tt = P.Tsum;
  const V3 gv{0, 0, g};
  const V3 a_hat = V3{P.alpha[0], P.alpha[1], P.alpha[2]} + mv(Jpa, dba) + mv(Jpg, dbg);
  const V3 b_hat = V3{P.beta[0], P.beta[1], P.beta[2]} + mv(Jva, dba) + mv(Jvg, dbg);
  const V3 theta = mv(Jqg, dbg);
  const Q gam{P.gamma[0], P.gamma[1], P.gamma[2], P.gamma[3]};
  const Q g_hat = qm(gam, qexp(theta));
  const V3 wp = (0.5 * Tt * Tt) * gv + tj - ti - Tt * vi, wv = Tt * gv + vj - vi;
  const V3 rp = mtv(Ri, wp) - a_hat, rv = mtv(Ri, wv) - b_hat;
  const Q Mq = qm(qc(gam), qm(qc(qi), qj)), E = qm(qc(g_hat), qm(qc(qi), qj));
  double r[15] = {rp.x, rp.y, rp.z, 2 * E.x, 2 * E.y, 2 * E.z, rv.x, rv.y, rv.z};
  for (int k = 0; k < 3; k++) { r[9 + k] = sbj[3 + k] - sbi[3 + k]; r[12 + k] = sbj[6 + k] - sbi[6 + k]; }
  for (int a = 0; a < 15; a++) { double v = 0; for (int m = a; m < 15; m++) v += P.W[15 * a + m] * r[m]; rw[a] = v; }
  if (!Jw) return;
  double Jr[450] = {0};
  const M3 Rit = tr(Ri), Gq = quat_rjac(E), GRjT = mm(Gq, tr(Rj)), dp_th = mm(Rit, skew(wp)), dv_th = mm(Rit, skew(wv));
  const M3 dq_bg = mm(mm(Gq, tr(q2R(Mq))), mm(so3_Jr({-theta.x, -theta.y, -theta.z}), Jqg));
  for (int a = 0; a < 3; a++)
    for (int b = 0; b < 3; b++) {
      const int ab = 3 * a + b;
      Jr[30 * a + b] = dp_th.m[ab]; Jr[30 * a + 3 + b] = -Rit.m[ab]; Jr[30 * a + 6 + b] = -Tt * Rit.m[ab];
      Jr[30 * a + 9 + b] = -Jpa.m[ab]; Jr[30 * a + 12 + b] = -Jpg.m[ab]; Jr[30 * a + 18 + b] = Rit.m[ab];
      Jr[30 * (3 + a) + b] = -GRjT.m[ab]; Jr[30 * (3 + a) + 12 + b] = -dq_bg.m[ab]; Jr[30 * (3 + a) + 15 + b] = GRjT.m[ab];
      Jr[30 * (6 + a) + b] = dv_th.m[ab]; Jr[30 * (6 + a) + 6 + b] = -Rit.m[ab]; Jr[30 * (6 + a) + 9 + b] = -Jva.m[ab];
      Jr[30 * (6 + a) + 12 + b] = -Jvg.m[ab]; Jr[30 * (6 + a) + 21 + b] = Rit.m[ab];
    }
  for (int a = 0; a < 3; a++) { Jr[30 * (9 + a) + 9 + a] = -1; Jr[30 * (9 + a) + 24 + a] = 1; Jr[30 * (12 + a) + 12 + a] = -1; Jr[30 * (12 + a) + 27 + a] = 1; }
  for (int a = 0; a < 15; a++)
    for (int c = 0; c < 30; c++) { double v = 0; for (int m = a; m < 15; m++) v += P.W[15 * a + m] * Jr[30 * m + c]; Jw[30 * a + c] = v; }
}

// ---------------------------------------------------------------------------------------------- the solver
using Clock = std::chrono::steady_clock;
inline double secs(Clock::time_point a, Clock::time_point b) { return std::chrono::duration<double>(b - a).count(); }

struct Solver {
  const bap_problem* p;
  bap_options o;
  int K, per, L_in = 0, n_obs = 0, n_imu = 0, n_edge = 0, n_c = 0, n_cp = 0, nt = 0, n_vec = 0;
  bool vo;
  double a2r, a2e, g = 9.81;
  std::vector<int> lm_of, okf, olm, lm_ptr, off_pose, off_sb;
  std::vector<double> ouv, osig;
  std::vector<double> pose, sb, lm, cpose, csb, clm;
  std::vector<Pre> pre;
  std::vector<int> fimu_i, fimu_j, fe_i, fe_j, fe_src;
  // linearisation
  std::vector<double> r_o, Jp_o, Jl_o, r_i, J_i, r_e, J_e;
  std::vector<double> scale, colsq, diag, gvec, grad, sgrad, gn, step, xsol;
  // blocks
  std::vector<double> W, Hll, Hinv, bl;
  // Schur lists (block → pairs)
  std::vector<int> sb_hi, sb_lo, sb_ptr, sp_a, sp_b;
  // tile storage of S / L (lower triangle incl. fill), tile_id[i*nt+j] or -1
  std::vector<int> tile_id;
  std::vector<std::vector<int>> rowcols;   // for tile row i: sorted column tiles j < i present in L
  std::vector<std::vector<int>> colrows;   // for tile column j: sorted row tiles i > j
  std::vector<double> St, S0;              // S0: J^T J camera part before damping/Schur (kept for re-damping)
  long n_tiles = 0;
  double factor_flops = 0;
  double ph[6] = {0, 0, 0, 0, 0, 0};

  int col(int kf, int c) const { return c < 6 ? off_pose[kf] + c : off_sb[kf] + (c - 6); }
  double& S_at(std::vector<double>& A, int r, int c) {   // r >= c required
    const int id = tile_id[(size_t)(r / T) * nt + c / T];
    return A[(size_t)id * T * T + (size_t)(r % T) * T + (c % T)];
  }
  void add_sym(std::vector<double>& A, int r, int c, double v) { if (r >= c) S_at(A, r, c) += v; else S_at(A, c, r) += v; }

  int setup();
  double evaluate(const std::vector<double>& ps, const std::vector<double>& sbs, const std::vector<double>& lms, bool jac);
  void build_blocks();
  bool solve_linear(double mu);
  double jv(const std::vector<double>& v, double* jvr);
};

int Solver::setup() {
  vo = o.visual_only != 0;
  per = vo ? 6 : 15;
  K = p->K;
  a2r = o.cauchy_reproj > 0 ? o.cauchy_reproj * o.cauchy_reproj : 0.0;
  a2e = o.cauchy_edge > 0 ? o.cauchy_edge * o.cauchy_edge : 0.0;
  pose.assign(p->pose, p->pose + 7 * (size_t)K);
  sb.assign(9 * (size_t)K, 0.0);
  if (p->speedbias) std::memcpy(sb.data(), p->speedbias, sizeof(double) * 9 * K);
  // landmarks with >= 2 usable observations
  lm_ptr.assign(1, 0);
  for (int l = 0; l < p->L; l++) {
    int cnt = 0;
    for (int ob = p->lm_obs_ptr[l]; ob < p->lm_obs_ptr[l + 1]; ob++) if (!(p->obs_skip && p->obs_skip[ob])) cnt++;
    if (cnt < 2) continue;
    const int c = (int)lm_of.size();
    lm_of.push_back(l);
    for (int k = 0; k < 3; k++) lm.push_back(p->lm[3 * (size_t)l + k]);
    for (int ob = p->lm_obs_ptr[l]; ob < p->lm_obs_ptr[l + 1]; ob++) {
      if (p->obs_skip && p->obs_skip[ob]) continue;
      okf.push_back(p->obs_kf[ob]); olm.push_back(c);
      ouv.push_back(p->obs_uv[2 * (size_t)ob]); ouv.push_back(p->obs_uv[2 * (size_t)ob + 1]);
      osig.push_back(p->obs_sigma[ob]);
    }
    lm_ptr.push_back((int)okf.size());
  }
  L_in = (int)lm_of.size(); n_obs = (int)okf.size();
  if (!vo) for (int f = 0; f < p->n_imu; f++) { fimu_i.push_back(p->imu_i[f]); fimu_j.push_back(p->imu_j[f]); }
  n_imu = (int)fimu_i.size();
  for (int e = 0; e < p->n_edge; e++) { fe_i.push_back(p->edge_i[e]); fe_j.push_back(p->edge_j[e]); fe_src.push_back(e); }
  n_edge = (int)fe_i.size();
  // ---- column layout ----
  off_pose.assign(K, 0); off_sb.assign(K, 0);
  int n_total = 0;
  if (!vo) {
    std::vector<int> parent(K);
    std::iota(parent.begin(), parent.end(), 0);
    std::function<int(int)> find = [&](int x) { while (parent[x] != x) { parent[x] = parent[parent[x]]; x = parent[x]; } return x; };
    for (int f = 0; f < n_imu; f++) { const int a = find(fimu_i[f]), b = find(fimu_j[f]); if (a != b) parent[std::max(a, b)] = std::min(a, b); }
    std::vector<int> csz(K, 0), roots;
    for (int k = 0; k < K; k++) csz[find(k)]++;
    int cur = 0;
    for (int r = 0; r < K; r++) {
      if (find(r) != r || csz[r] < 2) continue;
      roots.push_back(r);
      cur = (cur + T - 1) / T * T;
      for (int k = r; k < K; k++) if (find(k) == r) { off_sb[k] = cur; cur += 9; }
    }
    cur = (cur + T - 1) / T * T;
    for (int k = 0; k < K; k++) if (csz[find(k)] < 2) { off_sb[k] = cur; cur += 9; }
    cur = (cur + T - 1) / T * T;
    for (int r : roots) { cur = (cur + T - 1) / T * T; for (int k = r; k < K; k++) if (find(k) == r) { off_pose[k] = cur; cur += 6; } }
    cur = (cur + T - 1) / T * T;
    for (int k = 0; k < K; k++) if (csz[find(k)] < 2) { off_pose[k] = cur; cur += 6; }
    n_total = cur;
  } else {
    for (int k = 0; k < K; k++) off_pose[k] = 6 * k;
    n_total = 6 * K;
  }
  n_c = K * per;
  n_cp = (n_total + T - 1) / T * T;
  nt = n_cp / T;
  n_vec = n_cp + 3 * L_in;
  // ---- tile structure + symbolic fill ----
  std::vector<uint8_t> mask((size_t)nt * nt, 0);
  auto mark = [&](int a0, int al, int b0, int bl_) {
    for (int ta = a0 / T; ta <= (a0 + al - 1) / T; ta++)
      for (int tb = b0 / T; tb <= (b0 + bl_ - 1) / T; tb++) mask[(size_t)std::max(ta, tb) * nt + std::min(ta, tb)] = 1;
  };
  {
    std::vector<int> ts;
    for (int l = 0; l < L_in; l++) {
      ts.clear();
      for (int a = lm_ptr[l]; a < lm_ptr[l + 1]; a++) {
        const int o0 = off_pose[okf[a]];
        for (int t = o0 / T; t <= (o0 + 5) / T; t++) if (std::find(ts.begin(), ts.end(), t) == ts.end()) ts.push_back(t);
      }
      for (size_t x = 0; x < ts.size(); x++) for (size_t y = 0; y <= x; y++) mask[(size_t)std::max(ts[x], ts[y]) * nt + std::min(ts[x], ts[y])] = 1;
    }
  }
  for (int f = 0; f < n_imu; f++) {
    const int ij[2] = {fimu_i[f], fimu_j[f]};
    for (int a = 0; a < 2; a++) for (int b = 0; b < 2; b++) {
      mark(off_pose[ij[a]], 6, off_pose[ij[b]], 6); mark(off_pose[ij[a]], 6, off_sb[ij[b]], 9); mark(off_sb[ij[a]], 9, off_sb[ij[b]], 9);
    }
  }
  for (int e = 0; e < n_edge; e++) { mark(off_pose[fe_i[e]], 6, off_pose[fe_j[e]], 6); mark(off_pose[fe_i[e]], 6, off_pose[fe_i[e]], 6); mark(off_pose[fe_j[e]], 6, off_pose[fe_j[e]], 6); }
  for (int k = 0; k < nt; k++) mask[(size_t)k * nt + k] = 1;
  for (int k = 0; k < nt; k++) {   // right-looking clique fill
    std::vector<int> rows;
    for (int i = k + 1; i < nt; i++) if (mask[(size_t)i * nt + k]) rows.push_back(i);
    for (size_t a = 0; a < rows.size(); a++) for (size_t b = 0; b <= a; b++) mask[(size_t)rows[a] * nt + rows[b]] = 1;
  }
  tile_id.assign((size_t)nt * nt, -1);
  rowcols.assign(nt, {}); colrows.assign(nt, {});
  n_tiles = 0;
  for (int i = 0; i < nt; i++) for (int j = 0; j <= i; j++) if (mask[(size_t)i * nt + j]) {
    tile_id[(size_t)i * nt + j] = (int)n_tiles++;
    if (j < i) { rowcols[i].push_back(j); colrows[j].push_back(i); }
  }
  St.assign((size_t)n_tiles * T * T, 0.0); S0 = St;
  factor_flops = 0;
  for (int j = 0; j < nt; j++) { const double m = (double)colrows[j].size(); factor_flops += 2.0 * T * T * T * (m + m * (m + 1) / 2); }
  // ---- Schur (block → pair) lists ----
  {
    struct Pr { uint64_t key; int a, b; };
    std::vector<Pr> prs;
    for (int l = 0; l < L_in; l++)
      for (int a = lm_ptr[l]; a < lm_ptr[l + 1]; a++)
        for (int b = lm_ptr[l]; b <= a; b++) prs.push_back({((uint64_t)(uint32_t)okf[a] << 32) | (uint32_t)okf[b], a, b});
    std::stable_sort(prs.begin(), prs.end(), [](const Pr& x, const Pr& y) { return x.key < y.key; });
    for (size_t i = 0; i < prs.size(); i++) {
      if (i == 0 || prs[i].key != prs[i - 1].key) { sb_hi.push_back((int)(prs[i].key >> 32)); sb_lo.push_back((int)(prs[i].key & 0xffffffffu)); sb_ptr.push_back((int)i); }
      sp_a.push_back(prs[i].a); sp_b.push_back(prs[i].b);
    }
    sb_ptr.push_back((int)prs.size());
  }
  // ---- IMU preintegration at the initial biases of KF j ----
  pre.resize(n_imu);
  bool ok = true;
  if (n_imu) g = p->imu_noise[4];
#pragma omp parallel for schedule(dynamic, 8) reduction(&& : ok)
  for (int f = 0; f < n_imu; f++) {
    const int s = p->imu_ptr[f], e = p->imu_ptr[f + 1], j = fimu_j[f];
    ok = repropagate(p->imu_dt + s, p->imu_acc + 3 * (size_t)s, p->imu_gyr + 3 * (size_t)s, e - s, p->imu_acc0 + 3 * (size_t)f,
                     p->imu_gyr0 + 3 * (size_t)f, &sb[9 * (size_t)j + 3], &sb[9 * (size_t)j + 6], p->imu_noise, pre[f]) && ok;
  }
  if (!ok) return 4;
  r_o.resize(2 * (size_t)n_obs); Jp_o.resize(12 * (size_t)n_obs); Jl_o.resize(6 * (size_t)n_obs);
  r_i.resize(15 * (size_t)n_imu); J_i.resize(450 * (size_t)n_imu); r_e.resize(6 * (size_t)n_edge); J_e.resize(72 * (size_t)n_edge);
  W.resize(18 * (size_t)n_obs); Hll.resize(6 * (size_t)L_in); Hinv.resize(6 * (size_t)L_in); bl.resize(3 * (size_t)L_in);
  scale.assign(n_vec, 0.0);
  for (int k = 0; k < K; k++) {
    if (!p->pose_const[k]) for (int c = 0; c < 6; c++) scale[col(k, c)] = 1.0;
    if (!vo) for (int c = 6; c < 15; c++) scale[col(k, c)] = 1.0;
  }
  for (int i = n_cp; i < n_vec; i++) scale[i] = 1.0;
  for (auto* v : {&colsq, &diag, &gvec, &grad, &sgrad, &gn, &step, &xsol}) v->assign(n_vec, 0.0);
  return 0;
}

// residuals (+ corrected, Jacobi-scaled Jacobians); returns the cost
double Solver::evaluate(const std::vector<double>& ps, const std::vector<double>& sbs, const std::vector<double>& lms, bool jac) {
  double cost = 0;
#pragma omp parallel for schedule(static) reduction(+ : cost)
  for (int ob = 0; ob < n_obs; ob++) {
    const int k = okf[ob], l = olm[ob], cam = p->cam_of_kf ? p->cam_of_kf[k] : 0;
    double r[2], Jp[12], Jl[6];
    f_reproj(&ps[7 * (size_t)k], p->extr + 7 * (size_t)cam, p->intr + 4 * (size_t)cam, p->dist + 4 * (size_t)cam, &lms[3 * (size_t)l], ouv[2 * (size_t)ob],
             ouv[2 * (size_t)ob + 1], osig[ob], r, jac ? Jp : nullptr, jac ? Jl : nullptr);
    double sc, c;
    cauchy(r[0] * r[0] + r[1] * r[1], a2r, &sc, &c);
    cost += c;
    if (!jac) continue;
    r_o[2 * (size_t)ob] = r[0] * sc; r_o[2 * (size_t)ob + 1] = r[1] * sc;
    const double* sp = &scale[off_pose[k]];
    const double* sl = &scale[n_cp + 3 * (size_t)l];
    for (int a = 0; a < 2; a++) {
      for (int c2 = 0; c2 < 6; c2++) Jp_o[12 * (size_t)ob + 6 * a + c2] = Jp[6 * a + c2] * sc * sp[c2];
      for (int c2 = 0; c2 < 3; c2++) Jl_o[6 * (size_t)ob + 3 * a + c2] = Jl[3 * a + c2] * sc * sl[c2];
    }
  }
#pragma omp parallel for schedule(dynamic, 16) reduction(+ : cost)
  for (int f = 0; f < n_imu; f++) {
    const int i = fimu_i[f], j = fimu_j[f];
    double rw[15], Jw[450];
    f_imu(&ps[7 * (size_t)i], &sbs[9 * (size_t)i], &ps[7 * (size_t)j], &sbs[9 * (size_t)j], pre[f], g, rw, jac ? Jw : nullptr);
    double s = 0;
    for (int a = 0; a < 15; a++) s += rw[a] * rw[a];
    cost += 0.5 * s;
    if (!jac) continue;
    std::memcpy(&r_i[15 * (size_t)f], rw, sizeof(rw));
    for (int c = 0; c < 30; c++) {
      const double sc = scale[col(c < 15 ? i : j, c < 15 ? c : c - 15)];
      for (int a = 0; a < 15; a++) J_i[450 * (size_t)f + 30 * a + c] = Jw[30 * a + c] * sc;
    }
  }
#pragma omp parallel for schedule(dynamic, 64) reduction(+ : cost)
  for (int e = 0; e < n_edge; e++) {
    const int i = fe_i[e], j = fe_j[e], s0 = fe_src[e];
    double r[6], J[72];
    f_between(&ps[7 * (size_t)i], &ps[7 * (size_t)j], p->edge_q + 4 * (size_t)s0, p->edge_t + 3 * (size_t)s0, p->edge_sqrt_info + 36 * (size_t)s0, r, jac ? J : nullptr);
    double s = 0;
    for (int a = 0; a < 6; a++) s += r[a] * r[a];
    double sc, c;
    cauchy(s, (p->edge_robust && p->edge_robust[s0]) ? a2e : 0.0, &sc, &c);
    cost += c;
    if (!jac) continue;
    for (int a = 0; a < 6; a++) r_e[6 * (size_t)e + a] = r[a] * sc;
    for (int a = 0; a < 6; a++)
      for (int c2 = 0; c2 < 12; c2++) J_e[72 * (size_t)e + 12 * a + c2] = J[12 * a + c2] * sc * scale[off_pose[c2 < 6 ? i : j] + (c2 < 6 ? c2 : c2 - 6)];
  }
  return cost;
}

// J^T J blocks (camera part → S0 tiles; landmark part Hll, W), gradient gvec = J^T r, colsq = diag(J^T J)
void Solver::build_blocks() {
  std::fill(S0.begin(), S0.end(), 0.0);
  std::fill(gvec.begin(), gvec.end(), 0.0);
  // per keyframe pose block from the observations: thread-private accumulation over a by-keyframe pass
  std::vector<double> Hpp(21 * (size_t)K, 0.0), gp(6 * (size_t)K, 0.0);
  {
    const int nth = omp_get_max_threads();
    std::vector<std::vector<double>> th_H(nth), th_g(nth);
#pragma omp parallel
    {
      const int t = omp_get_thread_num();
      th_H[t].assign(21 * (size_t)K, 0.0); th_g[t].assign(6 * (size_t)K, 0.0);
      double* H = th_H[t].data();
      double* G = th_g[t].data();
#pragma omp for schedule(static)
      for (int ob = 0; ob < n_obs; ob++) {
        const int k = okf[ob];
        const double* J = &Jp_o[12 * (size_t)ob];
        const double r0 = r_o[2 * (size_t)ob], r1 = r_o[2 * (size_t)ob + 1];
        int idx = 0;
        for (int r = 0; r < 6; r++) for (int c = 0; c <= r; c++) H[21 * (size_t)k + idx++] += J[r] * J[c] + J[6 + r] * J[6 + c];
        for (int r = 0; r < 6; r++) G[6 * (size_t)k + r] += J[r] * r0 + J[6 + r] * r1;
        const double* Jl = &Jl_o[6 * (size_t)ob];
        double* w = &W[18 * (size_t)ob];
        for (int a = 0; a < 6; a++) for (int b = 0; b < 3; b++) w[3 * a + b] = J[a] * Jl[b] + J[6 + a] * Jl[3 + b];
      }
    }
    for (int t = 0; t < nth; t++) {
      for (size_t i = 0; i < Hpp.size(); i++) Hpp[i] += th_H[t][i];
      for (size_t i = 0; i < gp.size(); i++) gp[i] += th_g[t][i];
    }
  }
  for (int k = 0; k < K; k++) {
    int idx = 0;
    const int b = off_pose[k];
    for (int r = 0; r < 6; r++) for (int c = 0; c <= r; c++) S_at(S0, b + r, b + c) += Hpp[21 * (size_t)k + idx++];
    for (int r = 0; r < 6; r++) gvec[b + r] += gp[6 * (size_t)k + r];
  }
#pragma omp parallel for schedule(static)
  for (int l = 0; l < L_in; l++) {
    double h[6] = {0, 0, 0, 0, 0, 0}, b[3] = {0, 0, 0};
    for (int ob = lm_ptr[l]; ob < lm_ptr[l + 1]; ob++) {
      const double* J = &Jl_o[6 * (size_t)ob];
      const double r0 = r_o[2 * (size_t)ob], r1 = r_o[2 * (size_t)ob + 1];
      h[0] += J[0] * J[0] + J[3] * J[3]; h[1] += J[0] * J[1] + J[3] * J[4]; h[2] += J[0] * J[2] + J[3] * J[5];
      h[3] += J[1] * J[1] + J[4] * J[4]; h[4] += J[1] * J[2] + J[4] * J[5]; h[5] += J[2] * J[2] + J[5] * J[5];
      b[0] += J[0] * r0 + J[3] * r1; b[1] += J[1] * r0 + J[4] * r1; b[2] += J[2] * r0 + J[5] * r1;
    }
    for (int i = 0; i < 6; i++) Hll[6 * (size_t)l + i] = h[i];
    for (int i = 0; i < 3; i++) { bl[3 * (size_t)l + i] = b[i]; gvec[n_cp + 3 * (size_t)l + i] = b[i]; }
    colsq[n_cp + 3 * (size_t)l] = h[0]; colsq[n_cp + 3 * (size_t)l + 1] = h[3]; colsq[n_cp + 3 * (size_t)l + 2] = h[5];
  }
  // factors (few thousand): sequential scatter
  for (int f = 0; f < n_imu; f++) {
    const double* J = &J_i[450 * (size_t)f];
    const double* r = &r_i[15 * (size_t)f];
    const int kf[2] = {fimu_i[f], fimu_j[f]};
    int cols[30];
    for (int c = 0; c < 30; c++) cols[c] = col(kf[c / 15], c % 15);
    for (int a = 0; a < 30; a++) {
      double gs = 0;
      for (int m = 0; m < 15; m++) gs += J[30 * m + a] * r[m];
      gvec[cols[a]] += gs;
      for (int b = 0; b <= a; b++) {
        double s = 0;
        for (int m = 0; m < 15; m++) s += J[30 * m + a] * J[30 * m + b];
        add_sym(S0, cols[a], cols[b], s);
      }
    }
  }
  for (int e = 0; e < n_edge; e++) {
    const double* J = &J_e[72 * (size_t)e];
    const double* r = &r_e[6 * (size_t)e];
    int cols[12];
    for (int c = 0; c < 12; c++) cols[c] = off_pose[c < 6 ? fe_i[e] : fe_j[e]] + c % 6;
    for (int a = 0; a < 12; a++) {
      double gs = 0;
      for (int m = 0; m < 6; m++) gs += J[12 * m + a] * r[m];
      gvec[cols[a]] += gs;
      for (int b = 0; b <= a; b++) {
        double s = 0;
        for (int m = 0; m < 6; m++) s += J[12 * m + a] * J[12 * m + b];
        add_sym(S0, cols[a], cols[b], s);
      }
    }
  }
  for (int i = 0; i < n_cp; i++) colsq[i] = scale[i] != 0.0 ? S_at(S0, i, i) : 0.0;
}

inline double clampd(double v) { return std::min(std::max(v, 1e-6), 1e32); }

// x = (J^T J + mu diag^2)^-1 J^T r via landmark elimination; false when the factorisation breaks down
bool Solver::solve_linear(double mu) {
  auto t0 = Clock::now();
#pragma omp parallel for schedule(static)
  for (int l = 0; l < L_in; l++) {
    const double* h = &Hll[6 * (size_t)l];
    const double a = h[0] + mu * clampd(colsq[n_cp + 3 * (size_t)l]), b = h[1], c = h[2];
    const double d = h[3] + mu * clampd(colsq[n_cp + 3 * (size_t)l + 1]), e = h[4], f = h[5] + mu * clampd(colsq[n_cp + 3 * (size_t)l + 2]);
    const double A = d * f - e * e, B = c * e - b * f, C = b * e - c * d, id = 1.0 / (a * A + b * B + c * C);
    double* o2 = &Hinv[6 * (size_t)l];
    o2[0] = A * id; o2[1] = B * id; o2[2] = C * id; o2[3] = (a * f - c * c) * id; o2[4] = (b * c - a * e) * id; o2[5] = (a * d - b * b) * id;
  }
  St = S0;
  std::vector<double> gs(n_cp, 0.0);
  // Y = W Hll^-1 per observation (kept thread-local in the pair loop: recomputed from W on the fly)
  const int nb = (int)sb_hi.size();
#pragma omp parallel for schedule(dynamic, 64)
  for (int b = 0; b < nb; b++) {
    double acc[36] = {0};
    for (int q = sb_ptr[b]; q < sb_ptr[b + 1]; q++) {
      const int oa = sp_a[q], ob = sp_b[q];
      const double* h = &Hinv[6 * (size_t)olm[oa]];
      const double H[9] = {h[0], h[1], h[2], h[1], h[3], h[4], h[2], h[4], h[5]};
      const double* wa = &W[18 * (size_t)oa];
      const double* wb = &W[18 * (size_t)ob];
      double Y[18];
      for (int r = 0; r < 6; r++) for (int c = 0; c < 3; c++) Y[3 * r + c] = wa[3 * r] * H[c] + wa[3 * r + 1] * H[3 + c] + wa[3 * r + 2] * H[6 + c];
      for (int r = 0; r < 6; r++) for (int c = 0; c < 6; c++) acc[6 * r + c] += Y[3 * r] * wb[3 * c] + Y[3 * r + 1] * wb[3 * c + 1] + Y[3 * r + 2] * wb[3 * c + 2];
    }
    const int br = off_pose[sb_hi[b]], bc = off_pose[sb_lo[b]];
    // distinct blocks write distinct entries; a diagonal block (hi == lo) only its lower triangle
    for (int r = 0; r < 6; r++)
      for (int c = 0; c < 6; c++) {
        if (sb_hi[b] == sb_lo[b] && c > r) continue;
        if (br + r >= bc + c) S_at(St, br + r, bc + c) -= acc[6 * r + c]; else S_at(St, bc + c, br + r) -= acc[6 * r + c];
      }
  }
  // reduced gradient: gs = g_c - sum_obs Y b_l
  {
    std::vector<double> yb(n_cp, 0.0);
    for (int ob = 0; ob < n_obs; ob++) {
      const double* h = &Hinv[6 * (size_t)olm[ob]];
      const double* bb = &bl[3 * (size_t)olm[ob]];
      const double t[3] = {h[0] * bb[0] + h[1] * bb[1] + h[2] * bb[2], h[1] * bb[0] + h[3] * bb[1] + h[4] * bb[2], h[2] * bb[0] + h[4] * bb[1] + h[5] * bb[2]};
      const double* w = &W[18 * (size_t)ob];
      double* y = &yb[off_pose[okf[ob]]];
      for (int r = 0; r < 6; r++) y[r] += w[3 * r] * t[0] + w[3 * r + 1] * t[1] + w[3 * r + 2] * t[2];
    }
    for (int i = 0; i < n_cp; i++) {
      if (scale[i] != 0.0) { const double dg = diag[i]; S_at(St, i, i) += mu * dg * dg; gs[i] = gvec[i] - yb[i]; }
      else { S_at(St, i, i) = 1.0; gs[i] = 0.0; }
    }
  }
  auto t1 = Clock::now();
  ph[1] += secs(t0, t1);
  // ---- left-looking tile Cholesky ----
  bool ok = true;
  for (int j = 0; j < nt && ok; j++) {
    const std::vector<int>& rows = colrows[j];
    const int m = (int)rows.size();
#pragma omp parallel for schedule(dynamic, 1)
    for (int q = -1; q < m; q++) {
      const int i = q < 0 ? j : rows[q];
      double* C = &St[(size_t)tile_id[(size_t)i * nt + j] * T * T];
      const std::vector<int>& ci = rowcols[i];
      const std::vector<int>& cj = rowcols[j];
      size_t a = 0, b = 0;
      while (a < ci.size() && b < cj.size() && ci[a] < j && cj[b] < j) {
        if (ci[a] < cj[b]) a++;
        else if (ci[a] > cj[b]) b++;
        else {
          const int k = ci[a];
          const double* Lik = &St[(size_t)tile_id[(size_t)i * nt + k] * T * T];
          const double* Ljk = &St[(size_t)tile_id[(size_t)j * nt + k] * T * T];
          if (i == j) tile_syrk(Lik, C); else tile_gemm_nt(Lik, Ljk, C);
          a++; b++;
        }
      }
    }
    double* D = &St[(size_t)tile_id[(size_t)j * nt + j] * T * T];
    if (!tile_potrf(D)) { ok = false; break; }
#pragma omp parallel for schedule(dynamic, 1)
    for (int q = 0; q < m; q++) tile_trsm(D, &St[(size_t)tile_id[(size_t)rows[q] * nt + j] * T * T]);
  }
  auto t2 = Clock::now();
  ph[2] += secs(t1, t2);
  if (!ok) return false;
  // ---- forward / backward substitution ----
  std::vector<double>& x = xsol;
  for (int i = 0; i < n_cp; i++) x[i] = gs[i];
  for (int j = 0; j < nt; j++) {
    const double* D = &St[(size_t)tile_id[(size_t)j * nt + j] * T * T];
    double* xj = &x[(size_t)j * T];
    for (int r = 0; r < T; r++) { double s = xj[r]; for (int c = 0; c < r; c++) s -= D[r * T + c] * xj[c]; xj[r] = s / D[r * T + r]; }
    const std::vector<int>& rows = colrows[j];
#pragma omp parallel for schedule(static)
    for (int q = 0; q < (int)rows.size(); q++) {
      const double* A = &St[(size_t)tile_id[(size_t)rows[q] * nt + j] * T * T];
      double* xi = &x[(size_t)rows[q] * T];
      for (int r = 0; r < T; r++) { double s = 0; for (int c = 0; c < T; c++) s += A[r * T + c] * xj[c]; xi[r] -= s; }
    }
  }
  for (int j = nt - 1; j >= 0; j--) {
    double* xj = &x[(size_t)j * T];
    const std::vector<int>& rows = colrows[j];
    for (int q = 0; q < (int)rows.size(); q++) {
      const double* A = &St[(size_t)tile_id[(size_t)rows[q] * nt + j] * T * T];
      const double* xi = &x[(size_t)rows[q] * T];
      for (int r = 0; r < T; r++) { const double v = xi[r]; if (v != 0.0) for (int c = 0; c < T; c++) xj[c] -= A[r * T + c] * v; }
    }
    const double* D = &St[(size_t)tile_id[(size_t)j * nt + j] * T * T];
    for (int r = T - 1; r >= 0; r--) { double s = xj[r]; for (int c = r + 1; c < T; c++) s -= D[c * T + r] * xj[c]; xj[r] = s / D[r * T + r]; }
  }
  // landmark back-substitution
#pragma omp parallel for schedule(static)
  for (int l = 0; l < L_in; l++) {
    double t[3] = {bl[3 * (size_t)l], bl[3 * (size_t)l + 1], bl[3 * (size_t)l + 2]};
    for (int ob = lm_ptr[l]; ob < lm_ptr[l + 1]; ob++) {
      const double* w = &W[18 * (size_t)ob];
      const double* xc = &x[off_pose[okf[ob]]];
      for (int a = 0; a < 6; a++) { t[0] -= w[3 * a] * xc[a]; t[1] -= w[3 * a + 1] * xc[a]; t[2] -= w[3 * a + 2] * xc[a]; }
    }
    const double* h = &Hinv[6 * (size_t)l];
    x[n_cp + 3 * (size_t)l] = h[0] * t[0] + h[1] * t[1] + h[2] * t[2];
    x[n_cp + 3 * (size_t)l + 1] = h[1] * t[0] + h[3] * t[1] + h[4] * t[2];
    x[n_cp + 3 * (size_t)l + 2] = h[2] * t[0] + h[4] * t[1] + h[5] * t[2];
  }
  ph[3] += secs(t2, Clock::now());
  for (int i = 0; i < n_vec; i++) if (!std::isfinite(x[i])) return false;
  return true;
}

// |J v|^2 and (J v).r
double Solver::jv(const std::vector<double>& v, double* jvr) {
  double s0 = 0, s1 = 0;
#pragma omp parallel for schedule(static) reduction(+ : s0, s1)
  for (int ob = 0; ob < n_obs; ob++) {
    const double* Jp = &Jp_o[12 * (size_t)ob];
    const double* Jl = &Jl_o[6 * (size_t)ob];
    const double* vp = &v[off_pose[okf[ob]]];
    const double* vl = &v[n_cp + 3 * (size_t)olm[ob]];
    double j0 = 0, j1 = 0;
    for (int c = 0; c < 6; c++) { j0 += Jp[c] * vp[c]; j1 += Jp[6 + c] * vp[c]; }
    for (int c = 0; c < 3; c++) { j0 += Jl[c] * vl[c]; j1 += Jl[3 + c] * vl[c]; }
    s0 += j0 * j0 + j1 * j1;
    s1 += j0 * r_o[2 * (size_t)ob] + j1 * r_o[2 * (size_t)ob + 1];
  }
  for (int f = 0; f < n_imu; f++) {
    const double* J = &J_i[450 * (size_t)f];
    for (int a = 0; a < 15; a++) {
      double t = 0;
      for (int c = 0; c < 30; c++) t += J[30 * a + c] * v[col(c < 15 ? fimu_i[f] : fimu_j[f], c % 15)];
      s0 += t * t; s1 += t * r_i[15 * (size_t)f + a];
    }
  }
  for (int e = 0; e < n_edge; e++) {
    const double* J = &J_e[72 * (size_t)e];
    for (int a = 0; a < 6; a++) {
      double t = 0;
      for (int c = 0; c < 12; c++) t += J[12 * a + c] * v[off_pose[c < 6 ? fe_i[e] : fe_j[e]] + c % 6];
      s0 += t * t; s1 += t * r_e[6 * (size_t)e + a];
    }
  }
  *jvr = s1;
  return s0;
}

int run(const bap_problem* p, const bap_options* o, bap_result* res, std::vector<double>* norms_out) {
  if (o->threads > 0) omp_set_num_threads(o->threads);
  if (p_setthr) p_setthr(1);   // BLAS calls are issued from OpenMP threads
  Solver S;
  S.p = p; S.o = *o;
  auto t_begin = Clock::now();
  int rc = S.setup();
  if (rc) return rc;
  S.ph[5] = secs(t_begin, Clock::now());
  const int K = S.K, n_vec = S.n_vec, n_cp = S.n_cp;
  constexpr double MIN_MU = 1e-8, MAX_MU = 1.0, MU_INC = 10.0;
  double radius = 1e4, mu = 1e-8;
  std::vector<double> hist;
  std::vector<int> status;
  auto t0 = Clock::now();
  // iteration 0: Jacobi scaling from the unscaled Jacobian at x0, then the scaled linearisation
  S.evaluate(S.pose, S.sb, S.lm, true);
  S.build_blocks();
  for (int i = 0; i < n_vec; i++) if (S.scale[i] != 0.0) S.scale[i] = 1.0 / (1.0 + std::sqrt(S.colsq[i]));
  double cost = S.evaluate(S.pose, S.sb, S.lm, true);
  S.ph[0] += secs(t0, Clock::now());
  auto xnorm = [&](const std::vector<double>& ps, const std::vector<double>& sbs, const std::vector<double>& lms) {
    double v = 0;
    for (int k = 0; k < K; k++) if (!p->pose_const[k]) for (int c = 0; c < 7; c++) v += ps[7 * (size_t)k + c] * ps[7 * (size_t)k + c];
    if (!S.vo) for (double s : sbs) v += s * s;
    for (double s : lms) v += s * s;
    return std::sqrt(v);
  };
  double x_norm = xnorm(S.pose, S.sb, S.lm);
  hist.push_back(cost);
  int it = 0, invalid_run = 0, term = 0;
  bool reuse = false, have_blocks = false;
  double alpha = 0, gn2 = 0, gg = 0, g_gn = 0, dogleg_norm = 0;
  S.cpose = S.pose; S.csb = S.sb; S.clm = S.lm;
  bool first = true;
  while (it < o->max_iterations) {
    it++;
    bool solver_ok = true;
    if (!reuse) {
      reuse = true;
      auto ta = Clock::now();
      if (!have_blocks) { S.build_blocks(); have_blocks = true; }
      double gmax = 0;
      for (int i = 0; i < n_vec; i++) {
        const bool act = S.scale[i] != 0.0;
        S.diag[i] = act ? std::sqrt(clampd(S.colsq[i])) : 1.0;
        S.grad[i] = act ? S.gvec[i] / S.diag[i] : 0.0;
        S.sgrad[i] = S.grad[i] / S.diag[i];
        if (act) gmax = std::max(gmax, std::fabs(S.gvec[i] / S.scale[i]));
      }
      if (first && gmax <= 1e-10) { term = 1; it = 0; break; }
      first = false;
      double dummy;
      const double JgJg = S.jv(S.sgrad, &dummy);
      S.ph[1] += secs(ta, Clock::now());
      bool solved = false;
      while (mu < MAX_MU) {
        if (!S.solve_linear(mu)) { mu *= MU_INC; continue; }
        solved = true;
        break;
      }
      solver_ok = solved;
      if (solved) {
        gn2 = gg = g_gn = 0;
        for (int i = 0; i < n_vec; i++) {
          const double v = S.scale[i] != 0.0 ? -S.xsol[i] * S.diag[i] : 0.0;
          S.gn[i] = v;
          gn2 += v * v; gg += S.grad[i] * S.grad[i]; g_gn += S.grad[i] * v;
        }
        alpha = gg / JgJg;
      }
    }
    double model_change = -1.0;
    if (solver_ok) {
      auto ta = Clock::now();
      const double gn_norm = std::sqrt(gn2), g_norm = std::sqrt(gg);
      double ca, cb;
      if (gn_norm <= radius) { ca = 0; cb = 1; }
      else if (g_norm * alpha >= radius) { ca = -(radius / g_norm); cb = 0; }
      else {
        const double b_dot_a = -alpha * g_gn, a2 = (alpha * g_norm) * (alpha * g_norm), bma2 = a2 - 2 * b_dot_a + gn_norm * gn_norm, c = b_dot_a - a2;
        const double d = std::sqrt(c * c + bma2 * (radius * radius - a2));
        const double beta = c <= 0 ? (d - c) / bma2 : (radius * radius - a2) / (d + c);
        ca = -alpha * (1 - beta); cb = beta;
      }
      double dl2 = 0;
      for (int i = 0; i < n_vec; i++) {
        const double d = ca * S.grad[i] + cb * S.gn[i];
        dl2 += d * d;
        S.step[i] = S.scale[i] != 0.0 ? d / S.diag[i] : 0.0;
      }
      dogleg_norm = std::sqrt(dl2);
      double jvr;
      const double jv2 = S.jv(S.step, &jvr);
      model_change = -(jvr + 0.5 * jv2);
      if (model_change > 0) {
        invalid_run = 0;
        double step2 = 0, cand2 = 0;
        for (int k = 0; k < K; k++) {
          double d[6], out[7];
          const bool cst = p->pose_const[k] != 0;
          for (int c = 0; c < 6; c++) d[c] = cst ? 0.0 : S.step[S.off_pose[k] + c] * S.scale[S.off_pose[k] + c];
          if (cst) std::memcpy(out, &S.pose[7 * (size_t)k], 56); else pose_plus(&S.pose[7 * (size_t)k], d, out);
          for (int c = 0; c < 7; c++) { S.cpose[7 * (size_t)k + c] = out[c]; const double df = out[c] - S.pose[7 * (size_t)k + c]; step2 += df * df; if (!cst) cand2 += out[c] * out[c]; }
          for (int c = 0; c < 9; c++) {
            double v = S.sb[9 * (size_t)k + c];
            if (!S.vo) { const double dd = S.step[S.off_sb[k] + c] * S.scale[S.off_sb[k] + c]; v += dd; step2 += dd * dd; cand2 += v * v; }
            S.csb[9 * (size_t)k + c] = v;
          }
        }
        for (int i = 0; i < 3 * S.L_in; i++) { const double dd = S.step[n_cp + i] * S.scale[n_cp + i]; const double v = S.lm[i] + dd; S.clm[i] = v; step2 += dd * dd; cand2 += v * v; }
        const double ccost = S.evaluate(S.cpose, S.csb, S.clm, false);
        S.ph[4] += secs(ta, Clock::now());
        if (std::sqrt(step2) <= 1e-8 * (x_norm + 1e-8)) { term = 2; status.push_back(4); break; }
        if (std::fabs(cost - ccost) <= 1e-6 * cost) { term = 3; status.push_back(4); break; }
        const double rho = (cost - ccost) / model_change;
        if (rho > 1e-3) {
          S.pose.swap(S.cpose); S.sb.swap(S.csb); S.lm.swap(S.clm);
          cost = ccost; x_norm = std::sqrt(cand2);
          auto tb = Clock::now();
          S.evaluate(S.pose, S.sb, S.lm, true);
          have_blocks = false;
          S.ph[0] += secs(tb, Clock::now());
          if (rho < 0.25) radius *= 0.5;
          if (rho > 0.75) radius = std::max(radius, 3.0 * dogleg_norm);
          mu = std::max(MIN_MU, 2.0 * mu / MU_INC);
          reuse = false;
          status.push_back(1);
        } else {
          radius *= 0.5; reuse = true; status.push_back(2);
        }
        hist.push_back(cost);
        continue;
      }
      S.ph[4] += secs(ta, Clock::now());
    }
    invalid_run++;
    status.push_back(3);
    hist.push_back(cost);
    if (invalid_run > 5) { term = 4; break; }
    mu *= MU_INC; reuse = false;
  }
  // ---- results ----
  if (res->pose) std::memcpy(res->pose, S.pose.data(), sizeof(double) * 7 * K);
  if (res->speedbias) std::memcpy(res->speedbias, S.sb.data(), sizeof(double) * 9 * K);
  if (res->lm) {
    std::memcpy(res->lm, p->lm, sizeof(double) * 3 * (size_t)p->L);
    for (int c = 0; c < S.L_in; c++) std::memcpy(res->lm + 3 * (size_t)S.lm_of[c], &S.lm[3 * (size_t)c], 24);
  }
  res->iterations = it; res->termination = term; res->initial_cost = hist.front(); res->final_cost = cost;
  res->n_cost_history = 0;
  if (res->cost_history && res->cost_history_cap > 0) { const int n = std::min<int>(res->cost_history_cap, (int)hist.size()); for (int i = 0; i < n; i++) res->cost_history[i] = hist[i]; res->n_cost_history = n; }
  if (res->step_status && res->cost_history_cap > 0) { const int n = std::min<int>(res->cost_history_cap, (int)status.size()); for (int i = 0; i < n; i++) res->step_status[i] = (uint8_t)status[i]; }
  for (int i = 0; i < 6; i++) res->phase_s[i] = S.ph[i];
  res->factor_flops = S.factor_flops;
  if (norms_out) {   // loss-corrected reprojection residual norms at the final state (optimization_be.cpp:270-274)
    norms_out->assign(p->n_obs, -1.0);
    S.scale.assign(S.n_vec, 1.0);
    S.evaluate(S.pose, S.sb, S.lm, true);
    int c = 0;
    for (int l = 0; l < S.L_in; l++) {
      const int lo = S.lm_of[l];
      for (int ob = p->lm_obs_ptr[lo]; ob < p->lm_obs_ptr[lo + 1]; ob++) {
        if (p->obs_skip && p->obs_skip[ob]) continue;
        (*norms_out)[ob] = std::sqrt(S.r_o[2 * (size_t)c] * S.r_o[2 * (size_t)c] + S.r_o[2 * (size_t)c + 1] * S.r_o[2 * (size_t)c + 1]);
        c++;
      }
    }
  }
  return 0;
}

}  // namespace

// path of an OpenBLAS shared library exporting scipy_cblas_* / scipy_LAPACKE_dpotrf (the one bundled with scipy); returns 0
// when the four kernels were found (otherwise the plain C loops stay in use)
BAP_API int bap_init_blas(const char* path) {
  void* h = dlopen(path, RTLD_NOW | RTLD_LOCAL);
  if (!h) return 1;
  auto sym = [&](const char* a, const char* b) { void* s = dlsym(h, a); return s ? s : dlsym(h, b); };
  p_dgemm = (dgemm_t)sym("scipy_cblas_dgemm", "cblas_dgemm");
  p_dsyrk = (dsyrk_t)sym("scipy_cblas_dsyrk", "cblas_dsyrk");
  p_dtrsm = (dtrsm_t)sym("scipy_cblas_dtrsm", "cblas_dtrsm");
  p_dpotrf = (dpotrf_t)sym("scipy_LAPACKE_dpotrf", "LAPACKE_dpotrf");
  p_setthr = (setthr_t)sym("scipy_openblas_set_num_threads", "openblas_set_num_threads");
  if (!p_dgemm || !p_dsyrk || !p_dtrsm || !p_dpotrf) { p_dgemm = nullptr; p_dsyrk = nullptr; p_dtrsm = nullptr; p_dpotrf = nullptr; return 2; }
  return 0;
}

BAP_API int bap_solve(const bap_problem* p, const bap_options* o, bap_result* r) { return run(p, o, r, nullptr); }

// Optimization::GlobalBundleAdjustment (optimization_be.cpp:56-618): round 1 (5 iterations, loop edges without loss) +
// outlier purge on the loss-corrected residual norms (:270-290), round 2 from the map state with Cauchy(1) on the loops
BAP_API int bap_gba(const bap_problem* p, int iterations_limit, int visual_only, int outlier_removal, double th_outlier, int threads,
                    bap_result* r, uint8_t* obs_removed) {
  std::vector<uint8_t> skip(std::max(p->n_obs, 1), 0), rb0(std::max(p->n_edge, 1), 0), rb1(std::max(p->n_edge, 1), 1);
  if (p->obs_skip) std::memcpy(skip.data(), p->obs_skip, p->n_obs);
  bap_options o{5, visual_only, 1.0, 1.0, threads};
  if (outlier_removal) {
    bap_problem p1 = *p;
    p1.edge_robust = rb0.data();
    bap_result r1{};
    std::vector<double> norms;
    int rc = run(&p1, &o, &r1, &norms);
    if (rc) return rc;
    for (int i = 0; i < p->n_obs; i++) if (norms[i] > th_outlier) skip[i] = 1;
  }
  if (obs_removed) for (int i = 0; i < p->n_obs; i++) obs_removed[i] = skip[i] && !(p->obs_skip && p->obs_skip[i]);
  bap_problem p2 = *p;
  p2.obs_skip = skip.data();
  p2.edge_robust = rb1.data();
  o.max_iterations = iterations_limit;
  return run(&p2, &o, r, nullptr);
}

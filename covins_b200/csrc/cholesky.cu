// cholesky.cu — dense FP64 Cholesky factorisation + triangular solves of the reduced camera system (K8).
//
// Replaces the CHOLMOD factorisation inside Ceres' SPARSE_SCHUR (optimization_be.cpp:258,561,1025).  The
// reduced camera matrix S (n = 6K or 15K) is stored dense, row-major, lower triangle significant, padded to a
// multiple of the 128 tile.  At EuRoC scale every keyframe is covisible with hundreds of others (all agents fly
// the same hall), so S is ~10 % block-dense before fill and a dense tiled factorisation is the right shape for
// the GPU; this is the one BA stage that is a true GEMM and runs on the FP64 tensor cores (DMMA,
// mma.sync.m8n8k4.f64 — tcgen05 has no FP64 kind).
//
// Right-looking, panel width 128:
//   potrf_inv_kernel   1 CTA: factor the 128x128 diagonal tile in shared memory, write L, write L^-1
//   trsm_kernel        row tiles below: A(i,k) <- A(i,k) * Linv^T            (128^3 DMMA GEMM per CTA)
//   syrk_kernel        trailing tiles (i >= j > k): A(i,j) -= A(i,k) A(j,k)^T (128^3 DMMA GEMM per CTA)
// Solves use the stored tile inverses: forward L y = b, backward L^T x = y, one launch per tile column.
// All reductions have a fixed order → bit-reproducible run to run.
#include "cholesky.cuh"

namespace cvb_chol {

constexpr int KC = 32;        // K chunk staged in shared memory
constexpr int LDS = KC + 4;   // padded row stride (doubles): conflict-free m8n8k4 fragment loads
constexpr int GEMM_THREADS = 512;

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(cvb_smem_addr(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

// acc(128x128) = A(128xT) * B(128xT)^T, both operands row-major with K contiguous (leading dims lda, ldb).
// 16 warps in a 4x4 grid, 32x32 per warp = 4x4 m8n8k4 tiles; K staged in chunks of 32, double buffered.
__device__ __forceinline__ void gemm_abt_128(const double* __restrict__ A, size_t lda, const double* __restrict__ B,
                                             size_t ldb, double (&acc)[4][4][2], double* smem) {
  double* As[2] = {smem, smem + 2 * T * LDS};
  double* Bs[2] = {smem + T * LDS, smem + 3 * T * LDS};
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = warp >> 2, wn = warp & 3;
#pragma unroll
  for (int i = 0; i < 4; i++)
#pragma unroll
    for (int j = 0; j < 4; j++) acc[i][j][0] = acc[i][j][1] = 0.0;

  auto load_chunk = [&](int kc, int st) {
#pragma unroll
    for (int it = 0; it < 4; it++) {
      const int u = tid + it * GEMM_THREADS;  // 2048 16-byte units per operand
      const int r = u >> 4, seg = u & 15;
      cp_async16(As[st] + r * LDS + seg * 2, A + (size_t)r * lda + kc * KC + seg * 2);
      cp_async16(Bs[st] + r * LDS + seg * 2, B + (size_t)r * ldb + kc * KC + seg * 2);
    }
    cp_async_commit();
  };
  constexpr int NCH = T / KC;
  load_chunk(0, 0);
  for (int kc = 0; kc < NCH; kc++) {
    const int st = kc & 1;
    if (kc + 1 < NCH) {
      load_chunk(kc + 1, st ^ 1);
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    const double* a_base = As[st] + (wm * 32 + (lane >> 2)) * LDS + (lane & 3);
    const double* b_base = Bs[st] + (wn * 32 + (lane >> 2)) * LDS + (lane & 3);
#pragma unroll
    for (int ks = 0; ks < KC / 4; ks++) {
      double a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; i++) a[i] = a_base[i * 8 * LDS + ks * 4];
#pragma unroll
      for (int j = 0; j < 4; j++) b[j] = b_base[j * 8 * LDS + ks * 4];
#pragma unroll
      for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) dmma(acc[i][j][0], acc[i][j][1], a[i], b[j]);
    }
    __syncthreads();
  }
}

constexpr size_t kGemmSmem = (size_t)4 * T * LDS * sizeof(double);  // 147456 B

// A(i,k) <- A(i,k) * Linv_k^T for the structurally non-zero row tiles i of tile column k (rows[])
__global__ void __launch_bounds__(GEMM_THREADS, 1) trsm_kernel(double* __restrict__ S, size_t ld, int k,
                                                                const double* __restrict__ linv_k,
                                                                const int* __restrict__ rows) {
  extern __shared__ __align__(16) double smem_d[];
  const int i = rows[blockIdx.x];
  double* At = S + (size_t)i * T * ld + (size_t)k * T;
  double acc[4][4][2];
  gemm_abt_128(At, ld, linv_k, T, acc, smem_d);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wm = warp >> 2, wn = warp & 3;
#pragma unroll
  for (int a = 0; a < 4; a++)
#pragma unroll
    for (int b = 0; b < 4; b++) {
      const int r = wm * 32 + a * 8 + (lane >> 2), c = wn * 32 + b * 8 + (lane & 3) * 2;
      *reinterpret_cast<double2*>(At + (size_t)r * ld + c) = make_double2(acc[a][b][0], acc[a][b][1]);
    }
}

// A(i,j) -= A(i,k) A(j,k)^T for the tile pairs (i >= j) of column k's non-zero rows: pi[]/pj[] enumerate them.
__global__ void __launch_bounds__(GEMM_THREADS, 1) syrk_kernel(double* __restrict__ S, size_t ld, int k,
                                                                const int* __restrict__ pi, const int* __restrict__ pj) {
  extern __shared__ __align__(16) double smem_d[];
  const int i = pi[blockIdx.x], j = pj[blockIdx.x];
  const double* Ai = S + (size_t)i * T * ld + (size_t)k * T;
  const double* Aj = S + (size_t)j * T * ld + (size_t)k * T;
  double* C = S + (size_t)i * T * ld + (size_t)j * T;
  double acc[4][4][2];
  gemm_abt_128(Ai, ld, Aj, ld, acc, smem_d);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wm = warp >> 2, wn = warp & 3;
#pragma unroll
  for (int a = 0; a < 4; a++)
#pragma unroll
    for (int b = 0; b < 4; b++) {
      const int r = wm * 32 + a * 8 + (lane >> 2), c = wn * 32 + b * 8 + (lane & 3) * 2;
      double2* p = reinterpret_cast<double2*>(C + (size_t)r * ld + c);
      double2 v = *p;
      v.x -= acc[a][b][0];
      v.y -= acc[a][b][1];
      *p = v;
    }
}

// Factor the diagonal tile k in shared memory and invert the factor, one CTA of 512 threads.
//   Cholesky: blocked right-looking, 16-wide block columns.  Per block column: (a) the 16x16 diagonal block is factored
//   register-resident by one warp (rank-1 updates broadcast with shuffles) and inverted; (b) the panel below it is
//   multiplied by that inverse (no division chains); (c) the trailing sub-matrix gets its rank-16 update with 4x4
//   register tiles (0.5 shared-memory loads per multiply-add — the naive form is LDS-bandwidth bound).
//   Inverse: in place by recursive doubling (16 → 32 → 64 → 128): X21 = -C^-1 (B A^-1), 4x4 register tiles, a 32 KB
//   scratch tile; the 16x16 diagonal inverses come from the factorisation.
// L is written back to S, the inverse (row-major 128x128, zeros above the diagonal) to linv_k.  flag[0] |= 1 on a
// non-positive pivot (matrix not positive definite → the caller raises mu, as Ceres does on LINEAR_SOLVER_FAILURE).
constexpr int LDP = T + 1;
constexpr int PB = 16;
constexpr int POTRF_THREADS = 512;
constexpr size_t kPotrfSmem = ((size_t)T * LDP + (size_t)64 * 64 + (size_t)(T / PB) * PB * PB) * sizeof(double);

// C[4][4] (+)= sum_m P[i][m] * Q[j][m], rows of P/Q at stride ldp/ldq, m in [m0, m1)
__device__ __forceinline__ void tile4x4(const double* __restrict__ P, int ldp, const double* __restrict__ Q, int ldq, int m0,
                                        int m1, double (&c)[4][4]) {
  for (int m = m0; m < m1; m++) {
    double pv[4], qv[4];
#pragma unroll
    for (int i = 0; i < 4; i++) { pv[i] = P[i * ldp + m]; qv[i] = Q[i * ldq + m]; }
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
      for (int j = 0; j < 4; j++) c[i][j] += pv[i] * qv[j];
  }
}

__global__ void __launch_bounds__(POTRF_THREADS, 1) potrf_inv_kernel(double* __restrict__ S, size_t ld, int k,
                                                                     double* __restrict__ linv_k, int* __restrict__ flag) {
  extern __shared__ __align__(16) double smem_d[];
  double* a = smem_d;                     // [T][LDP]
  double* tmp = smem_d + T * LDP;         // 64 x 64 scratch
  double* binv = tmp + 64 * 64;           // [8][16][16] inverses of the diagonal 16x16 blocks of L
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  double* At = S + (size_t)k * T * ld + (size_t)k * T;
  for (int u = tid; u < T * T; u += POTRF_THREADS) {
    const int r = u / T, c = u % T;
    a[r * LDP + c] = (c <= r) ? At[(size_t)r * ld + c] : 0.0;
  }
  __syncthreads();
  // ---------------- Cholesky ----------------
  for (int jb = 0; jb < T / PB; jb++) {
    const int c0 = jb * PB, nbelow = T - c0 - PB;
    if (warp == 0) {
      // (a) 16x16 diagonal block: lane i (< 16; lanes 16-31 mirror) owns row i
      const int i = lane & 15;
      double row[PB];
#pragma unroll
      for (int c = 0; c < PB; c++) row[c] = a[(c0 + i) * LDP + c0 + c];
      double rd[PB];   // reciprocal diagonal (identical on all lanes)
#pragma unroll
      for (int j = 0; j < PB; j++) {
        double piv = __shfl_sync(0xffffffffu, row[j], j);
        if (!(piv > 0.0)) {
          if (lane == 0) atomicOr(flag, 1);
          piv = 1.0;
        }
        const double rinv = rsqrt(piv);
        rd[j] = rinv;
        const double lij = (i >= j) ? row[j] * rinv : 0.0;   // column j of L
        row[j] = lij;
#pragma unroll
        for (int c = j + 1; c < PB; c++) {
          const double lcj = __shfl_sync(0xffffffffu, lij, c);
          if (i >= c) row[c] -= lij * lcj;
        }
      }
      if (lane < PB) {
#pragma unroll
        for (int c = 0; c < PB; c++) a[(c0 + i) * LDP + c0 + c] = (c <= i) ? row[c] : 0.0;
      }
      __syncwarp();
      // inverse of the block: lane j (< 16) builds column j by forward substitution (L read back from shared memory)
      if (lane < PB) {
        const int j = lane;
        double x[PB];
#pragma unroll
        for (int r = 0; r < PB; r++) {
          double sacc = (r == j) ? 1.0 : 0.0;
#pragma unroll
          for (int m = 0; m < PB; m++)
            if (m < r) sacc -= a[(c0 + r) * LDP + c0 + m] * ((m >= j) ? x[m] : 0.0);
          x[r] = (r >= j) ? sacc * rd[r] : 0.0;
        }
#pragma unroll
        for (int r = 0; r < PB; r++) binv[(jb * PB + r) * PB + j] = x[r];   // Linv_block[r][j]
      }
    }
    __syncthreads();
    if (nbelow > 0) {
      // (b) panel: X[r][j] = sum_{m <= j} A[r][c0+m] * Linv_block[j][m]   (X Ljj^T = A)
      for (int u = tid; u < nbelow * PB; u += POTRF_THREADS) {
        const int r = c0 + PB + u / PB, j = u % PB;
        const double* ar = a + r * LDP + c0;
        const double* bj = binv + (jb * PB + j) * PB;
        double sacc = 0.0;
#pragma unroll
        for (int m = 0; m < PB; m++)
          if (m <= j) sacc += ar[m] * bj[m];
        tmp[u] = sacc;
      }
      __syncthreads();
      for (int u = tid; u < nbelow * PB; u += POTRF_THREADS) a[(c0 + PB + u / PB) * LDP + c0 + (u % PB)] = tmp[u];
      __syncthreads();
      // (c) trailing update A[r][c] -= sum_{m<16} L[r][c0+m] L[c][c0+m] for r >= c >= c0+16, 4x4 register tiles
      const int nt4 = nbelow / 4, ntiles = nt4 * (nt4 + 1) / 2;
      for (int u = tid; u < ntiles; u += POTRF_THREADS) {
        int tr = (int)((sqrtf(8.0f * (float)u + 1.0f) - 1.0f) * 0.5f);
        while ((tr + 1) * (tr + 2) / 2 <= u) tr++;
        while (tr * (tr + 1) / 2 > u) tr--;
        const int tc = u - tr * (tr + 1) / 2;
        const int r0 = c0 + PB + 4 * tr, cc0 = c0 + PB + 4 * tc;
        double c4[4][4];
#pragma unroll
        for (int i = 0; i < 4; i++)
#pragma unroll
          for (int j = 0; j < 4; j++) c4[i][j] = 0.0;
        tile4x4(a + r0 * LDP + c0, LDP, a + cc0 * LDP + c0, LDP, 0, PB, c4);
#pragma unroll
        for (int i = 0; i < 4; i++)
#pragma unroll
          for (int j = 0; j < 4; j++)
            if (cc0 + j <= r0 + i) a[(r0 + i) * LDP + cc0 + j] -= c4[i][j];
      }
      __syncthreads();
    }
  }
  for (int u = tid; u < T * T; u += POTRF_THREADS) {
    const int r = u / T, c = u % T;
    if (c <= r) At[(size_t)r * ld + c] = a[r * LDP + c];
  }
  __syncthreads();
  // ---------------- inverse of L, in place ----------------
  // level 0: the diagonal 16x16 blocks were inverted during the factorisation
  for (int u = tid; u < (T / PB) * PB * PB; u += POTRF_THREADS) {
    const int blk = u / (PB * PB), r = (u / PB) % PB, c = u % PB;
    a[(blk * PB + r) * LDP + blk * PB + c] = binv[u];
  }
  __syncthreads();
  // levels h = 16, 32, 64: for each pair (A = inv at [p,p], C = inv at [p+h,p+h], B at [p+h,p]):
  //   tmp = B * A  (A lower triangular),  B <- -C * tmp  (C lower triangular); 4x4 register tiles
  for (int h = PB; h < T; h *= 2) {
    const int npairs = T / (2 * h), h4 = h / 4;
    for (int u = tid; u < npairs * h4 * h4; u += POTRF_THREADS) {
      const int pr = u / (h4 * h4), r0 = 4 * ((u / h4) % h4), cc0 = 4 * (u % h4);
      const int p0 = pr * 2 * h;
      double c4[4][4];
#pragma unroll
      for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) c4[i][j] = 0.0;
      // (B A)[r][c] = sum_m B[r][m] A[m][c], A[m][c] = 0 for m < c
      for (int m = cc0; m < h; m++) {
        double bv[4], av[4];
#pragma unroll
        for (int i = 0; i < 4; i++) { bv[i] = a[(p0 + h + r0 + i) * LDP + p0 + m]; av[i] = a[(p0 + m) * LDP + p0 + cc0 + i]; }
#pragma unroll
        for (int i = 0; i < 4; i++)
#pragma unroll
          for (int j = 0; j < 4; j++) c4[i][j] += bv[i] * av[j];
      }
#pragma unroll
      for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) tmp[pr * h * h + (r0 + i) * h + cc0 + j] = c4[i][j];
    }
    __syncthreads();
    for (int u = tid; u < npairs * h4 * h4; u += POTRF_THREADS) {
      const int pr = u / (h4 * h4), r0 = 4 * ((u / h4) % h4), cc0 = 4 * (u % h4);
      const int p0 = pr * 2 * h;
      double c4[4][4];
#pragma unroll
      for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) c4[i][j] = 0.0;
      // (C tmp)[r][c] = sum_{m <= r} C[r][m] tmp[m][c]
      for (int m = 0; m < r0 + 4; m++) {
        double cv[4], tv[4];
#pragma unroll
        for (int i = 0; i < 4; i++) { cv[i] = a[(p0 + h + r0 + i) * LDP + p0 + h + m]; tv[i] = tmp[pr * h * h + m * h + cc0 + i]; }
#pragma unroll
        for (int i = 0; i < 4; i++)
#pragma unroll
          for (int j = 0; j < 4; j++) c4[i][j] += cv[i] * tv[j];
      }
#pragma unroll
      for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) a[(p0 + h + r0 + i) * LDP + p0 + cc0 + j] = -c4[i][j];
    }
    __syncthreads();
  }
  for (int u = tid; u < T * T; u += POTRF_THREADS) {
    const int r = u / T, c = u % T;
    linv_k[u] = (c <= r) ? a[r * LDP + c] : 0.0;
  }
}

// Triangular solves.  One launch per tile column; every CTA recomputes the tiny diagonal product (128x128 mat-vec with
// the stored tile inverse) and then updates its own off-diagonal tile.  512 threads, all loads of a mat-vec are issued
// before the first use (32 independent coalesced loads per thread) — these kernels are latency-, not bandwidth-bound.
constexpr int SOLVE_THREADS = 512;

// out[r] = sum_c A[r][c] x[c]   (A row-major 128x128 with leading dimension lda; lanes run along c)
__device__ __forceinline__ void matvec_rows(const double* __restrict__ A, size_t lda, const double* x, double* out) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;   // 16 warps x 8 rows
  double v[8][4];
#pragma unroll
  for (int i = 0; i < 8; i++)
#pragma unroll
    for (int j = 0; j < 4; j++) v[i][j] = A[(size_t)(warp * 8 + i) * lda + lane + 32 * j];
#pragma unroll
  for (int i = 0; i < 8; i++) {
    double s = (v[i][0] * x[lane] + v[i][1] * x[lane + 32]) + (v[i][2] * x[lane + 64] + v[i][3] * x[lane + 96]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) out[warp * 8 + i] = s;
  }
}
// out[c] = sum_m A[m][c] x[m]   (threads run along c; 4 groups of 128 threads split m, partials combined in order)
__device__ __forceinline__ void matvec_cols(const double* __restrict__ A, size_t lda, const double* x, double* out,
                                            double* part /*[4][128]*/) {
  const int c = threadIdx.x & 127, g = threadIdx.x >> 7;
  double v[32];
#pragma unroll
  for (int i = 0; i < 32; i++) v[i] = A[(size_t)(g * 32 + i) * lda + c];
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < 32; i++) s += v[i] * x[g * 32 + i];
  part[g * T + c] = s;
  __syncthreads();
  if (threadIdx.x < T) out[c] = (part[c] + part[T + c]) + (part[2 * T + c] + part[3 * T + c]);
  __syncthreads();
}

// forward step k: y_k = Linv_k b_k (written by CTA 0), b_i -= L(i,k) y_k for the non-zero row tiles i > k (rows[])
__global__ void __launch_bounds__(SOLVE_THREADS) fwd_kernel(const double* __restrict__ L, size_t ld, int k,
                                                            const double* __restrict__ linv, double* __restrict__ b,
                                                            double* __restrict__ y, const int* __restrict__ rows) {
  __shared__ double bk[T], yk[T], upd[T];
  const int tid = threadIdx.x;
  if (tid < T) bk[tid] = b[(size_t)k * T + tid];
  __syncthreads();
  matvec_rows(linv + (size_t)k * T * T, T, bk, yk);   // Linv is stored with zeros above the diagonal
  __syncthreads();
  if (blockIdx.x == 0) {
    if (tid < T) y[(size_t)k * T + tid] = yk[tid];
    return;
  }
  const int i = rows[blockIdx.x - 1];
  matvec_rows(L + (size_t)i * T * ld + (size_t)k * T, ld, yk, upd);
  __syncthreads();
  if (tid < T) b[(size_t)i * T + tid] -= upd[tid];
}

// backward step k: x_k = Linv_k^T y_k (CTA 0), y_i -= L(k,i)^T x_k for the non-zero column tiles i < k of row k (cols[])
__global__ void __launch_bounds__(SOLVE_THREADS) bwd_kernel(const double* __restrict__ L, size_t ld, int k,
                                                            const double* __restrict__ linv, double* __restrict__ y,
                                                            double* __restrict__ x, const int* __restrict__ cols) {
  __shared__ double ykk[T], xk[T], upd[T], part[4 * T];
  const int tid = threadIdx.x;
  if (tid < T) ykk[tid] = y[(size_t)k * T + tid];
  __syncthreads();
  matvec_cols(linv + (size_t)k * T * T, T, ykk, xk, part);
  if (blockIdx.x == 0) {
    if (tid < T) x[(size_t)k * T + tid] = xk[tid];
    return;
  }
  const int i = cols[blockIdx.x - 1];
  matvec_cols(L + (size_t)k * T * ld + (size_t)i * T, ld, xk, upd, part);
  if (tid < T) y[(size_t)i * T + tid] -= upd[tid];
}

// Symbolic phase (host): tile-level structure of L from the tile-level structure of S (lower, nt x nt, row-major
// bools, diagonal forced).  Right-looking elimination: the non-zero rows of column k become a clique.
void TilePlan::build(int nt_, std::vector<uint8_t> mask) {
  nt = nt_;
  h_col_ptr.assign(1, 0); h_row_idx.clear(); h_pair_ptr.assign(1, 0); h_pair_i.clear(); h_pair_j.clear(); h_pair_split.clear();
  for (int k = 0; k < nt; k++) mask[(size_t)k * nt + k] = 1;
  for (int k = 0; k < nt; k++) {
    std::vector<int> rows;
    for (int i = k + 1; i < nt; i++)
      if (mask[(size_t)i * nt + k]) rows.push_back(i);
    // pairs whose column tile is k+1 first ("panel" part: all the next step's potrf/trsm depend on), then the rest
    int n_a = 0;
    for (int pass = 0; pass < 2; pass++)
      for (size_t a = 0; a < rows.size(); a++)
        for (size_t b = 0; b <= a; b++) {
          const bool is_a = rows[b] == k + 1;
          if ((pass == 0) != is_a) continue;
          mask[(size_t)rows[a] * nt + rows[b]] = 1;
          h_pair_i.push_back(rows[a]);
          h_pair_j.push_back(rows[b]);
          if (is_a) n_a++;
        }
    h_pair_split.push_back(n_a);
    h_row_idx.insert(h_row_idx.end(), rows.begin(), rows.end());
    h_col_ptr.push_back((int)h_row_idx.size());
    h_pair_ptr.push_back((int)h_pair_i.size());
  }
  h_rowc_ptr.assign(1, 0); h_rowc_idx.clear();
  for (int k = 0; k < nt; k++) {
    for (int i = 0; i < k; i++)
      if (mask[(size_t)k * nt + i]) h_rowc_idx.push_back(i);
    h_rowc_ptr.push_back((int)h_rowc_idx.size());
  }
  n_tiles_L = (long)h_row_idx.size() + nt;
  // flops actually executed: one 128^3 GEMM (2 flop per MAC) per trsm tile and per syrk pair
  flops = 2.0 * T * T * T * ((double)h_row_idx.size() + (double)h_pair_i.size());
}

int TilePlan::upload(cvb_ctx* ctx, cudaStream_t st) {
  release();
  auto up = [&](int** d, const std::vector<int>& h) -> int {
    const size_t n = h.size() ? h.size() : 1;
    CVB_CUDA(ctx, cudaMalloc(d, n * sizeof(int)));
    if (h.size()) CVB_CUDA(ctx, cudaMemcpyAsync(*d, h.data(), h.size() * sizeof(int), cudaMemcpyHostToDevice, st));
    return CVB_OK;
  };
  int rc;
  if ((rc = up(&d_row_idx, h_row_idx)) || (rc = up(&d_pair_i, h_pair_i)) || (rc = up(&d_pair_j, h_pair_j)) ||
      (rc = up(&d_rowc_idx, h_rowc_idx)))
    return rc;
  CVB_CUDA(ctx, cudaStreamSynchronize(st));
  return CVB_OK;
}

void TilePlan::release() {
  if (d_row_idx) cudaFree(d_row_idx);
  if (d_pair_i) cudaFree(d_pair_i);
  if (d_pair_j) cudaFree(d_pair_j);
  if (d_rowc_idx) cudaFree(d_rowc_idx);
  d_row_idx = d_pair_i = d_pair_j = d_rowc_idx = nullptr;
}

int factor(cvb_ctx* ctx, double* S, int n_pad, double* linv, int* d_flag, const TilePlan& plan, cudaStream_t st,
           const FactorStreams* fs) {
  static bool attr = false;
  if (!attr) {
    CVB_CUDA(ctx, cudaFuncSetAttribute(trsm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kGemmSmem));
    CVB_CUDA(ctx, cudaFuncSetAttribute(syrk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kGemmSmem));
    CVB_CUDA(ctx, cudaFuncSetAttribute(potrf_inv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kPotrfSmem));
    attr = true;
  }
  const int nt = n_pad / T;
  CVB_REQUIRE(ctx, plan.nt == nt, "tile plan does not match the matrix");
  CVB_CUDA(ctx, cudaMemsetAsync(d_flag, 0, sizeof(int), st));
  cudaStream_t st2 = fs ? fs->bulk : nullptr;
  cudaEvent_t* ev = fs ? fs->ev : nullptr;
  const int n_gs = (fs && !plan.h_col_group.empty()) ? fs->n_group : 0;
  // Lookahead (depth 1) when a second stream is given: the diagonal-tile kernel and the panel solve of step k+1 only
  // need the "panel" part of step k's trailing update (pairs in tile column k+1); the bulk of the update runs on the
  // second (low-priority) stream concurrently.  ev[2k] = panel solve of step k done, ev[2k+1] = bulk update done.
  // Independent column groups (the IMU chains of different agents, see TilePlan::h_col_group) run on their own
  // streams: their tile columns are pure latency chains (diagonal tile → panel → tiny update) that do not share tiles.
  const bool la = st2 != nullptr && ev != nullptr;
  int last_bulk = -1;
  bool forked = false;
  std::vector<char> used(n_gs > 0 ? n_gs : 1, 0);
  for (int k = 0; k < nt; k++) {
    const int grp = n_gs > 0 ? plan.h_col_group[k] : -1;
    cudaStream_t s = st;
    if (grp >= 0) {
      if (!forked) {
        CVB_CUDA(ctx, cudaEventRecord(fs->fork, st));
        forked = true;
      }
      s = fs->group[grp % n_gs];
      if (!used[grp % n_gs]) {
        CVB_CUDA(ctx, cudaStreamWaitEvent(s, fs->fork, 0));
        used[grp % n_gs] = 1;
      }
    } else if (forked) {   // first column after the grouped ones: join
      for (int g = 0; g < n_gs; g++)
        if (used[g]) {
          CVB_CUDA(ctx, cudaEventRecord(fs->join[g], fs->group[g]));
          CVB_CUDA(ctx, cudaStreamWaitEvent(st, fs->join[g], 0));
          used[g] = 0;
        }
      forked = false;
    }
    potrf_inv_kernel<<<1, POTRF_THREADS, kPotrfSmem, s>>>(S, (size_t)n_pad, k, linv + (size_t)k * T * T, d_flag);
    CVB_CHECK_LAUNCH(ctx);
    const int m = plan.h_col_ptr[k + 1] - plan.h_col_ptr[k];
    if (m > 0) {
      trsm_kernel<<<m, GEMM_THREADS, kGemmSmem, s>>>(S, (size_t)n_pad, k, linv + (size_t)k * T * T,
                                                     plan.d_row_idx + plan.h_col_ptr[k]);
      CVB_CHECK_LAUNCH(ctx);
      const bool la_k = la && grp < 0;
      const int p0 = plan.h_pair_ptr[k], np = plan.h_pair_ptr[k + 1] - p0, na = la_k ? plan.h_pair_split[k] : np;
      if (la_k) {
        CVB_CUDA(ctx, cudaEventRecord(ev[2 * k], s));
        if (last_bulk >= 0) CVB_CUDA(ctx, cudaStreamWaitEvent(s, ev[2 * last_bulk + 1], 0));
      }
      if (na > 0) {
        syrk_kernel<<<na, GEMM_THREADS, kGemmSmem, s>>>(S, (size_t)n_pad, k, plan.d_pair_i + p0, plan.d_pair_j + p0);
        CVB_CHECK_LAUNCH(ctx);
      }
      if (la_k && np - na > 0) {
        CVB_CUDA(ctx, cudaStreamWaitEvent(st2, ev[2 * k], 0));
        syrk_kernel<<<np - na, GEMM_THREADS, kGemmSmem, st2>>>(S, (size_t)n_pad, k, plan.d_pair_i + p0 + na,
                                                               plan.d_pair_j + p0 + na);
        CVB_CHECK_LAUNCH(ctx);
        CVB_CUDA(ctx, cudaEventRecord(ev[2 * k + 1], st2));
        last_bulk = k;
      }
    }
  }
  if (forked)
    for (int g = 0; g < n_gs; g++)
      if (used[g]) {
        CVB_CUDA(ctx, cudaEventRecord(fs->join[g], fs->group[g]));
        CVB_CUDA(ctx, cudaStreamWaitEvent(st, fs->join[g], 0));
      }
  if (la && last_bulk >= 0) CVB_CUDA(ctx, cudaStreamWaitEvent(st, ev[2 * last_bulk + 1], 0));   // join
  return CVB_OK;
}

// solves L L^T x = b; b is destroyed, tmp is scratch (n_pad), result in x
int solve(cvb_ctx* ctx, const double* L, int n_pad, const double* linv, double* b, double* tmp, double* x,
          const TilePlan& plan, cudaStream_t st) {
  const int nt = n_pad / T;
  for (int k = 0; k < nt; k++) {
    const int m = plan.h_col_ptr[k + 1] - plan.h_col_ptr[k];
    fwd_kernel<<<1 + m, SOLVE_THREADS, 0, st>>>(L, (size_t)n_pad, k, linv, b, tmp, plan.d_row_idx + plan.h_col_ptr[k]);
    CVB_CHECK_LAUNCH(ctx);
  }
  for (int k = nt - 1; k >= 0; k--) {
    const int m = plan.h_rowc_ptr[k + 1] - plan.h_rowc_ptr[k];
    bwd_kernel<<<1 + m, SOLVE_THREADS, 0, st>>>(L, (size_t)n_pad, k, linv, tmp, x, plan.d_rowc_idx + plan.h_rowc_ptr[k]);
    CVB_CHECK_LAUNCH(ctx);
  }
  return CVB_OK;
}

}  // namespace cvb_chol

// ---- test/diagnostic entry: solve A x = b for a host SPD matrix (row-major n x n) with the BA factorisation ----
extern "C" int cvb_dense_cholesky_solve(cvb_ctx* ctx, const double* A, int n, const double* b, double* x,
                                        double* factor_ms) {
  if (!ctx || !A || !b || !x || n <= 0) return CVB_ERR_INVALID;
  using namespace cvb_chol;
  const int np = ((n + T - 1) / T) * T;
  cudaStream_t st = ctx->stream;
  double* dS = (double*)cvb_ws(ctx, WS_T, (size_t)np * np * sizeof(double));
  double* dl = (double*)cvb_ws(ctx, WS_Q, (size_t)np * T * sizeof(double));
  double* dv = (double*)cvb_ws(ctx, WS_OUT0, (size_t)np * 3 * sizeof(double));
  int* dflag = (int*)cvb_ws(ctx, WS_FLAG, 16);
  if (!dS || !dl || !dv || !dflag) return CVB_ERR_CUDA;
  std::vector<double> hs((size_t)np * np, 0.0), hb(np, 0.0);
  for (int i = 0; i < n; i++)
    for (int j = 0; j <= i; j++) hs[(size_t)i * np + j] = A[(size_t)i * n + j];
  for (int i = n; i < np; i++) hs[(size_t)i * np + i] = 1.0;
  for (int i = 0; i < n; i++) hb[i] = b[i];
  CVB_CUDA(ctx, cudaMemcpyAsync(dS, hs.data(), hs.size() * sizeof(double), cudaMemcpyHostToDevice, st));
  CVB_CUDA(ctx, cudaMemcpyAsync(dv, hb.data(), np * sizeof(double), cudaMemcpyHostToDevice, st));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  // tile structure of the input (zero tiles are skipped) → symbolic fill
  const int nt = np / T;
  std::vector<uint8_t> mask((size_t)nt * nt, 0);
  for (int i = 0; i < n; i++)
    for (int j = 0; j <= i; j++)
      if (hs[(size_t)i * np + j] != 0.0) mask[(size_t)(i / T) * nt + (j / T)] = 1;
  TilePlan plan;
  plan.build(nt, mask);
  int rc = plan.upload(ctx, st);
  if (rc) return rc;
  cudaEventRecord(e0, st);
  rc = factor(ctx, dS, np, dl, dflag, plan, st, nullptr);
  cudaEventRecord(e1, st);
  if (rc) return rc;
  rc = solve(ctx, dS, np, dl, dv, dv + np, dv + 2 * np, plan, st);
  if (rc) return rc;
  int flag = 0;
  CVB_CUDA(ctx, cudaMemcpyAsync(&flag, dflag, sizeof(int), cudaMemcpyDeviceToHost, st));
  CVB_CUDA(ctx, cudaMemcpyAsync(hb.data(), dv + 2 * np, np * sizeof(double), cudaMemcpyDeviceToHost, st));
  CVB_CUDA(ctx, cudaStreamSynchronize(st));
  float ms = 0.f;
  cudaEventElapsedTime(&ms, e0, e1);
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  if (factor_ms) *factor_ms = ms;
  plan.release();
  if (flag) return cvb_fail(ctx, CVB_ERR_NUMERIC, "matrix is not positive definite");
  for (int i = 0; i < n; i++) x[i] = hb[i];
  return CVB_OK;
}

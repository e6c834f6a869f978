// pqrfact / psvdfact tails (reference: src/pqr.jl:297-305,330-340; src/psvd.jl:243-269,301-308;
// R*V at src/id.jl:153-158,354-360).
//
// B200 design (not the reference's geqrt + gesdd):
//  * QR of the skeleton columns C = A[:,sk] is a RANDOMISED-PRECONDITIONED CholeskyQR2: the sketch already
//    produced R11 with Omega*C = Q_B*R11, so Y = C*R11^{-1} is well conditioned (kappa = O(10..100)) even
//    though kappa(C) ~ 1/rtol; two Cholesky-QR passes on Y give Q orthonormal to machine precision and
//    R1 = R_y2 * R_y1 * R11.  All O(m k^2) work is GEMM-shaped and runs on the TMA + DMMA kernel.
//  * psvd: W = R1 [I T] P' is never formed.  Z = [I; T'] (n x k) is well conditioned, Z = Q_z R_z by
//    CholeskyQR2, so W = (R1 R_z') Q_z' P' and the ill-conditioning lives in the k x k core M = R1 R_z'
//    only, whose SVD is a one-sided Jacobi (Hestenes) on the graded matrix M' (accurate for small sigma).
#include "common.cuh"
#include "qrcp_exchange.cuh"
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <numeric>

namespace {

constexpr int TB = 32;

// C[:, i] = A[:, idx[i]-1]  (trans 'n', C is m x k)   or   C[:, i] = A[idx[i]-1, :]'  (trans 'c', C is n x k)
__global__ void gather_cols_kernel(char trans, const double* __restrict__ A, int64_t lda, int64_t mC, int64_t k,
                                   const int64_t* __restrict__ idx1, double* __restrict__ C, int64_t ldc) {
  for (int64_t i = blockIdx.x; i < k; i += gridDim.x) {
    const int64_t s = idx1[i] - 1;
    double* d = C + i * ldc;
    if (trans == 'n') {
      const double* a = A + s * lda;
      for (int64_t r = threadIdx.x; r < mC; r += blockDim.x) d[r] = a[r];
    } else {
      for (int64_t r = threadIdx.x; r < mC; r += blockDim.x) d[r] = A[s + r * lda];
    }
  }
}

__global__ void transpose2_kernel(const double* __restrict__ src, int64_t lds, int64_t rows, int64_t cols,
                                  double* __restrict__ dst, int64_t ldd) {
  __shared__ double t[32][33];
  const int64_t r0 = (int64_t)blockIdx.x * 32, c0 = (int64_t)blockIdx.y * 32;
  for (int y = threadIdx.y; y < 32; y += blockDim.y) {
    int64_t r = r0 + threadIdx.x, c = c0 + y;
    t[y][threadIdx.x] = (r < rows && c < cols) ? src[r + c * lds] : 0.0;
  }
  __syncthreads();
  for (int y = threadIdx.y; y < 32; y += blockDim.y) {
    int64_t c = c0 + threadIdx.x, r = r0 + y;
    if (r < rows && c < cols) dst[c + r * ldd] = t[threadIdx.x][y];
  }
}

// Y <- Y R^{-1}, R upper triangular k x k.  CTA = 64 rows; column blocks left to right.
__global__ void __launch_bounds__(256) trsolve_right_upper_kernel(int64_t rows, int k, const double* __restrict__ R,
                                                                  int64_t ldr, double* __restrict__ Y, int64_t ldy) {
  __shared__ double Ys[64][TB + 1];    // Ys[r][c] = Y[row0 + r, ib*32 + c]
  __shared__ double Rs[TB][TB + 1];    // Rs[r][c] = R[ib*32 + r, jb*32 + c]
  const int tid = threadIdx.x;
  const int tx = tid & 63;             // row within the panel
  const int ty = tid >> 6;             // 4 column groups of 8
  const int64_t row0 = (int64_t)blockIdx.x * 64;
  const int nblk = (k + TB - 1) / TB;
  for (int jb = 0; jb < nblk; ++jb) {
    const int c0 = jb * TB;
    double acc[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int c = c0 + ty * 8 + u;
      acc[u] = (row0 + tx < rows && c < k) ? Y[row0 + tx + (int64_t)c * ldy] : 0.0;
    }
    for (int ib = 0; ib < jb; ++ib) {
      __syncthreads();
      for (int e = tid; e < 64 * TB; e += 256) {
        const int rr = e & 63, cc = e >> 6;
        const int c = ib * TB + cc;
        Ys[rr][cc] = (row0 + rr < rows && c < k) ? Y[row0 + rr + (int64_t)c * ldy] : 0.0;
      }
      for (int e = tid; e < TB * TB; e += 256) {
        const int rr = e & 31, cc = e >> 5;
        const int r = ib * TB + rr, c = c0 + cc;
        Rs[rr][cc] = (r < k && c < k) ? R[r + (int64_t)c * ldr] : 0.0;
      }
      __syncthreads();
#pragma unroll 8
      for (int kk = 0; kk < TB; ++kk) {
        const double y = Ys[tx][kk];
#pragma unroll
        for (int u = 0; u < 8; ++u) acc[u] = fma(-y, Rs[kk][ty * 8 + u], acc[u]);
      }
    }
    __syncthreads();
    for (int e = tid; e < TB * TB; e += 256) {
      const int rr = e & 31, cc = e >> 5;
      const int r = c0 + rr, c = c0 + cc;
      Rs[rr][cc] = (r < k && c < k) ? R[r + (int64_t)c * ldr] : (rr == cc ? 1.0 : 0.0);
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) Ys[tx][ty * 8 + u] = acc[u];
    __syncthreads();
    if (ty == 0) {
      // forward substitution along the 32 columns of this block, one row per thread
      double y[TB];
#pragma unroll
      for (int c = 0; c < TB; ++c) y[c] = Ys[tx][c];
#pragma unroll
      for (int c = 0; c < TB; ++c) {
        y[c] = y[c] / Rs[c][c];
#pragma unroll
        for (int cc = c + 1; cc < TB; ++cc) y[cc] = fma(-y[c], Rs[c][cc], y[cc]);
      }
      if (row0 + tx < rows) {
#pragma unroll
        for (int c = 0; c < TB; ++c)
          if (c0 + c < k) Y[row0 + tx + (int64_t)(c0 + c) * ldy] = y[c];
      }
    }
    __syncthreads();
  }
}

// ---- blocked Cholesky G = L L^T, k x k; the factor is written TRANSPOSED (R = L^T, upper) into a separate matrix ----
// Per 32-column panel j two launches:
//  chol_diag_kernel   one warp: lane r keeps row r of the diagonal block in registers and factors it with shuffles
//                     (no barriers, reciprocal square roots), then inverts L_jj by column-oriented substitution; L_jj
//                     goes to R, L_jj^{-1} to a 32 x 32 scratch.
//  chol_trail_kernel  one CTA per trailing tile (I >= J): forms the two panel blocks it needs itself,
//                     L_I = P_I L_jj^{-T} and L_J = P_J L_jj^{-T} (small products, no substitution chain, no separate
//                     panel-solve launch), updates G[I,J] -= L_I L_J^T, and the J = 0 column of CTAs stores L_I into R.
//                     The unsolved panel blocks P stay untouched in G (nobody reads them after this launch), so there
//                     is no read/write race between CTAs.
// (one warp; G and Dinv are read / written by different CTAs across the panels of the fused kernel: L2 loads)
__device__ __forceinline__ void chol_diag_body(int k, int j0, const double* G, int64_t ldg, double* R, int64_t ldr,
                                               double* Dinv, int* info) {
  const int lane = threadIdx.x & 31;
  const int jbsz = min(TB, k - j0);
  double a[TB];      // row `lane` of the diagonal block (lower triangle meaningful)
#pragma unroll
  for (int c = 0; c < TB; ++c)
    a[c] = (lane < jbsz && c < jbsz) ? __ldcg(G + (j0 + lane) + (int64_t)(j0 + c) * ldg) : (lane == c ? 1.0 : 0.0);
  // Factorization and inversion run in ONE loop over the columns: as soon as column c of L exists, unknown c of the
  // substitution L v = e_j (lane j = column j of L^{-1}) is resolved and eliminated from the rows below.  The two
  // instruction streams are independent within a step, which is what hides the shuffle / FP64 latencies of a lone warp.
  double v[TB];
#pragma unroll
  for (int r = 0; r < TB; ++r) v[r] = (r == lane) ? 1.0 : 0.0;
#pragma unroll
  for (int c = 0; c < TB; ++c) {
    double d = __shfl_sync(0xffffffffu, a[c], c);
    if (!(d > 0.0)) {
      if (lane == 0 && c < jbsz) atomicExch(info, j0 + c + 1);
      d = 1.0;
    }
    const double rs = rsqrt(d);
    const double l = (lane == c) ? d * rs : a[c] * rs;              // column c of L (rows >= c meaningful)
    a[c] = l;
    v[c] *= rs;                                                     // 1 / L[c][c]
#pragma unroll
    for (int cc = c + 1; cc < TB; ++cc) {
      const double lcc = __shfl_sync(0xffffffffu, l, cc);           // L[cc][c]
      a[cc] = fma(-l, lcc, a[cc]);
      v[cc] = fma(-lcc, v[c], v[cc]);                                // the same L[cc][c] serves both streams
    }
  }
  if (lane < jbsz) {
#pragma unroll
    for (int c = 0; c < TB; ++c)
      if (c <= lane) R[(j0 + c) + (int64_t)(j0 + lane) * ldr] = a[c];      // R = L^T
  }
#pragma unroll
  for (int r = 0; r < TB; ++r) Dinv[r + lane * TB] = v[r];        // Dinv[r][j] = (L^{-1})[r][j], column-major 32 x 32
}

struct CholSmem {
  double Pi[TB][TB + 1];   // P_I, then reused
  double Pj[TB][TB + 1];
  double Di[TB][TB + 1];   // L_jj^{-1}
  double Li[TB][TB + 1];
  double Lj[TB][TB + 1];
};
__device__ __forceinline__ void chol_trail_body(CholSmem& sm, int tile, int k, int j0, double* G, int64_t ldg, double* R,
                                                int64_t ldr, const double* Dinv) {
  auto& Pi = sm.Pi;
  auto& Pj = sm.Pj;
  auto& Di = sm.Di;
  auto& Li = sm.Li;
  auto& Lj = sm.Lj;
  const int tid = threadIdx.x;
  // linear index -> (I, J), I >= J
  int t = tile, I = 0;
  while (t > I) {
    t -= I + 1;
    ++I;
  }
  const int J = t;
  const int ri = j0 + (I + 1) * TB, rj = j0 + (J + 1) * TB;
  for (int e = tid; e < TB * TB; e += 256) {
    const int rr = e & 31, cc = e >> 5;
    Pi[rr][cc] = (ri + rr < k && j0 + cc < k) ? __ldcg(G + (ri + rr) + (int64_t)(j0 + cc) * ldg) : 0.0;
    Pj[rr][cc] = (rj + rr < k && j0 + cc < k) ? __ldcg(G + (rj + rr) + (int64_t)(j0 + cc) * ldg) : 0.0;
    Di[rr][cc] = __ldcg(Dinv + rr + cc * TB);
  }
  __syncthreads();
  const int tx = tid & 31, ty = tid >> 5;
  // L_I[r][c] = sum_{t <= c} P_I[r][t] * Linv[c][t]
  {
    double ai[4] = {0.0, 0.0, 0.0, 0.0}, aj[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll 8
    for (int kk = 0; kk < TB; ++kk) {
      const double pi = Pi[tx][kk], pj = Pj[tx][kk];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const double dv = Di[ty * 4 + u][kk];       // zero above the diagonal (kk > c)
        ai[u] = fma(pi, dv, ai[u]);
        aj[u] = fma(pj, dv, aj[u]);
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      Li[tx][ty * 4 + u] = ai[u];
      Lj[tx][ty * 4 + u] = aj[u];
      if (J == 0 && ri + tx < k && j0 + ty * 4 + u < k) R[(j0 + ty * 4 + u) + (int64_t)(ri + tx) * ldr] = ai[u];
    }
  }
  __syncthreads();
  {
    double acc[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll 8
    for (int kk = 0; kk < TB; ++kk) {
      const double x = Li[tx][kk];
#pragma unroll
      for (int u = 0; u < 4; ++u) acc[u] = fma(x, Lj[ty * 4 + u][kk], acc[u]);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int c = ty * 4 + u;
      if (ri + tx < k && rj + c < k) {
        double* gp = G + (ri + tx) + (int64_t)(rj + c) * ldg;
        *gp = __ldcg(gp) - acc[u];
      }
    }
  }
}

// The whole blocked Cholesky in ONE cooperative launch (round 1: two launches per 32-column panel, 31 at k = 497):
// per panel CTA 0 factors and inverts the diagonal block, a grid barrier, every CTA updates its share of the trailing
// tiles, a grid barrier.  `bar` is zeroed before the launch; a lost CTA makes the others time out and flag `info`.
__device__ __forceinline__ bool chol_grid_barrier(unsigned* bar, unsigned target, int* s_fail) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(bar, 1u);
    unsigned v, spins = 0;
    do {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(bar) : "memory");
    } while (v < target && ++spins < (1u << 22));
    if (v < target) *s_fail = 1;
    __threadfence();
  }
  __syncthreads();
  return *s_fail == 0;
}

__global__ void __launch_bounds__(256, 1) chol_fused_kernel(int k, double* G, int64_t ldg, double* R, int64_t ldr, double* Dinv,
                                                            int* info, unsigned* bar) {
  __shared__ CholSmem sm;
  __shared__ int s_fail;
  if (threadIdx.x == 0) s_fail = 0;
  const int nblk = (k + TB - 1) / TB, Gn = gridDim.x;
  unsigned epoch = 0;
  for (int jb = 0; jb < nblk; ++jb) {
    const int j0 = jb * TB;
    if (blockIdx.x == 0 && threadIdx.x < 32) chol_diag_body(k, j0, G, ldg, R, ldr, Dinv, info);
    const int nb = nblk - jb - 1;
    if (nb == 0) break;
    if (!chol_grid_barrier(bar, ++epoch * Gn, &s_fail)) break;
    const int nt = nb * (nb + 1) / 2;
    for (int t = blockIdx.x; t < nt; t += Gn) {
      __syncthreads();
      chol_trail_body(sm, t, k, j0, G, ldg, R, ldr, Dinv);
    }
    if (!chol_grid_barrier(bar, ++epoch * Gn, &s_fail)) break;
  }
  if (s_fail && threadIdx.x == 0) atomicExch(info, -1);
}

// ---- one-sided Jacobi (Hestenes) on the columns of X (k x k), rotations accumulated into J ----
// Block-cyclic and persistent.  The k columns form nblk blocks of BC columns; in every outer step each CTA owns one
// block PAIR (chess-tournament schedule over the blocks), stages its 2*BC columns of X and of J in shared memory and
// runs one full round-robin sweep of plane rotations among them there.  A TEAM of TS threads works on one column
// pair (thread e holds rows e, e+TS, ... in registers for the dot products and the rotation), so a panel round
// costs a few hundred cycles instead of a warp-serial pass over the columns.
// Synchronisation between outer steps is POINT TO POINT: block b carries a step counter in global memory; the CTA
// that needs (bI, bJ) at step g waits until both counters reached g, and bumps them after writing the blocks back.
// There is no grid-wide barrier on the step path (one per sweep only, for the convergence vote), so a CTA whose
// pairs have converged runs ahead instead of waiting for the slowest.
// The rotations themselves are the classical ones: computed from freshly accumulated a = |x_p|^2, b = |x_q|^2,
// c = x_p.x_q of the CURRENT columns (no Gram-matrix shortcut), so small singular values keep their accuracy.
struct JacobiParams {
  int k, nblk;
  double* X;
  int64_t ldx;
  double* J;
  int64_t ldj;
  double tol;
  int max_sweeps;
  unsigned* bar;       // grid barrier counter (zeroed before launch)
  unsigned* bstep;     // [nblk] outer steps completed per block (zeroed before launch)
  int* rotated;        // [max_sweeps] "a rotation happened in this sweep"
  int* out;            // [0] sweeps done, [1] converged
  LL32* mbox;          // [nblk][mbox_words] hand-over mailboxes: self-validating 32-byte words (two doubles + the step stamp)
  size_t mbox_words;   // words per block: BC * kp / 2 (x2 with the J panel)
};

__device__ __forceinline__ void grid_barrier(unsigned* bar, unsigned target) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(bar, 1u);
    unsigned v;
    do {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(bar) : "memory");
    } while (v < target);
    __threadfence();
  }
  __syncthreads();
}

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc)
               : "memory");
}

// ---- bulk (TMA 1-D) copies for the block hand-over: one elected thread moves a whole column per instruction ----
__device__ __forceinline__ uint32_t jsmem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void jmbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(jsmem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void jmbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(jsmem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void jmbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(jsmem_u32(bar)), "r"(parity)
        : "memory");
  } while (!ok);
}
__device__ __forceinline__ void bulk_g2s(void* sdst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(jsmem_u32(sdst)),
               "l"(gsrc), "r"(bytes), "r"(jsmem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* gdst, const void* ssrc, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(jsmem_u32(ssrc)), "r"(bytes)
               : "memory");
}

__device__ __forceinline__ void rr_pair(int np, int stage, int i, int& p, int& q) {
  // round-robin (chess tournament) over np = even number of players: np-1 stages of np/2 disjoint pairs
  if (i == 0) {
    p = np - 1;
    q = stage % (np - 1);
  } else {
    p = (stage + i) % (np - 1);
    q = (stage + np - 1 - i) % (np - 1);
  }
  if (p > q) {
    const int t = p;
    p = q;
    q = t;
  }
}

template <int TS>
__device__ __forceinline__ void team_sync(int team) {
  if (TS == 32) __syncwarp();
  else asm volatile("bar.sync %0, %1;" ::"r"(team + 1), "r"(TS) : "memory");
}

// One plane rotation of the column pair (p, q) of the panel by a team of TS threads (thread e: rows e, e+TS, ...).
// Returns true if the pair was rotated.  cs/sn come from ONE warp of the team (the FP64 pipe is the scarce unit:
// 64 DFMA/clk/SM), written with two reciprocal square roots and no division:
//   cos(2 theta) = |d| / h,  sin(2 theta) = sign(d) 2c / h,  h = sqrt(d^2 + 4c^2),  d = b - a
//   cs = sqrt((1 + cos 2theta) / 2),  sn = sin(2 theta) / (2 cs)         (|theta| <= pi/4)
template <int TS, int R, bool WJ>
__device__ __forceinline__ bool rotate_pair(double* __restrict__ Xs, double* __restrict__ Js, int kp, int k, int p, int q,
                                            int team, int e, double* __restrict__ rb, double tol2, double thr2, int* s_big) {
  constexpr int WPT = TS / 32;
  const int lane = e & 31, tw = e >> 5;
  double* xp = Xs + (size_t)p * kp;
  double* xq = Xs + (size_t)q * kp;
  double u[R], v[R];
  double a = 0.0, b = 0.0, c = 0.0, a1 = 0.0, b1 = 0.0, c1 = 0.0;      // two chains per sum: DFMA latency ~ 16 cycles
#pragma unroll
  for (int i = 0; i < R; ++i) {
    const int r = e + TS * i;
    u[i] = (r < k) ? xp[r] : 0.0;
    v[i] = (r < k) ? xq[r] : 0.0;
    if (i & 1) {
      a1 = fma(u[i], u[i], a1);
      b1 = fma(v[i], v[i], b1);
      c1 = fma(u[i], v[i], c1);
    } else {
      a = fma(u[i], u[i], a);
      b = fma(v[i], v[i], b);
      c = fma(u[i], v[i], c);
    }
  }
  a += a1;
  b += b1;
  c += c1;
  // warp reduction of the three sums with 6 exchanges instead of 15: after the xor-16 stage a lane keeps (a, c) or
  // (b, -), after the xor-8 stage one value; lane 0 ends with a, lane 8 with c, lane 16 with b.
  {
    const bool hi16 = lane & 16, hi8 = lane & 8;
    const double s0 = hi16 ? a : b, k0 = hi16 ? b : a;          // send the one I do not own, keep the other
    double w0 = k0 + __shfl_xor_sync(0xffffffffu, s0, 16);      // lanes < 16: a, lanes >= 16: b
    double w1 = c + __shfl_xor_sync(0xffffffffu, c, 16);        // c (both halves hold the same partial pair sum)
    // xor 8: lanes with bit 3 clear keep w0, the others keep w1 (c only matters in the low half)
    const double s1 = hi8 ? w0 : w1, k1 = hi8 ? w1 : w0;
    double w = k1 + __shfl_xor_sync(0xffffffffu, s1, 8);
    w += __shfl_xor_sync(0xffffffffu, w, 4);
    w += __shfl_xor_sync(0xffffffffu, w, 2);
    w += __shfl_xor_sync(0xffffffffu, w, 1);
    if (WPT > 1) {
      // rb: [WPT][4] partial sums, then [2] = (cs, sn) (sn == 2.0: no rotation)
      if ((lane & 7) == 0 && lane < 24) rb[tw * 4 + (lane == 0 ? 0 : (lane == 16 ? 1 : 2))] = w;
      team_sync<TS>(team);
    } else {
      a = __shfl_sync(0xffffffffu, w, 0);
      c = __shfl_sync(0xffffffffu, w, 8);
      b = __shfl_sync(0xffffffffu, w, 16);
    }
  }
  if (tw == (team % WPT)) {
    if (WPT > 1) {
      a = b = c = 0.0;
#pragma unroll
      for (int w = 0; w < WPT; ++w) {
        a += rb[w * 4 + 0];
        b += rb[w * 4 + 1];
        c += rb[w * 4 + 2];
      }
    }
    double cs = 1.0, sn = 2.0;
    const double ab = a * b, cc2 = c * c;
    if (cc2 > thr2 * ab && lane == 0) *s_big = 1;      // a rotation that is NOT tiny (see the convergence vote)
    if (cc2 > tol2 * ab) {               // converged pair: |c| <= tol * sqrt(a b)
      const double d = b - a, c2 = 2.0 * c;
      const double rh = rsqrt(fma(d, d, c2 * c2));
      const double hc = fma(0.5 * fabs(d), rh, 0.5);          // (1 + cos 2theta) / 2  in [1/2, 1]
      const double rc = rsqrt(hc);
      cs = hc * rc;
      sn = (d >= 0.0 ? 0.5 : -0.5) * (c2 * rh) * rc;
    }
    if (lane == 0) {
      rb[WPT * 4 + 0] = cs;
      rb[WPT * 4 + 1] = sn;
    }
  }
  team_sync<TS>(team);
  const double cs = rb[WPT * 4 + 0], sn = rb[WPT * 4 + 1];
  if (sn == 2.0) return false;
  if (WJ) {
    double* jp = Js + (size_t)p * kp;
    double* jq = Js + (size_t)q * kp;
#pragma unroll
    for (int i = 0; i < R; ++i) {
      const int r = e + TS * i;
      if (r < k) {
        const double ju = jp[r], jv = jq[r];
        xp[r] = fma(cs, u[i], -sn * v[i]);
        xq[r] = fma(sn, u[i], cs * v[i]);
        jp[r] = fma(cs, ju, -sn * jv);
        jq[r] = fma(sn, ju, cs * jv);
      }
    }
  } else {
#pragma unroll
    for (int i = 0; i < R; ++i) {
      const int r = e + TS * i;
      if (r < k) {
        xp[r] = fma(cs, u[i], -sn * v[i]);
        xq[r] = fma(sn, u[i], cs * v[i]);
      }
    }
  }
  return true;
}

// BC = column pairs per panel (= columns per block), TS = threads per pair, R = rows per thread (k <= TS * R)
//
// Ordering: ODD-EVEN TRANSPOSITION over the nblk block positions (Luk & Park): even steps pair positions (2i, 2i+1),
// odd steps (2i+1, 2i+2); after every step the two blocks of a pair swap positions.  nblk steps make a sweep in which
// every pair of blocks meets exactly once.  CTA i owns pair i of the current step, so between steps it KEEPS one of
// its two blocks in shared memory and exchanges only the other one with a neighbour (half the traffic of a
// round-robin tournament, where both blocks move every step).
// WJ = false: the rotations are not accumulated (no J panel in shared memory, half the exchange traffic); the caller
// recovers J from the triangular starting matrix, J = X0^{-1} X_final (Drmac-Veselic).
template <int BC, int TS, int R, bool WJ>
__global__ void __launch_bounds__(BC * TS, 1) jacobi_team_kernel(JacobiParams P) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int NC = 2 * BC;                       // columns in a panel
  constexpr int WPT = TS / 32;                     // warps per team
  constexpr int RBS = WPT * 4 + 2;                 // doubles of reduction scratch per team and parity
  const int k = P.k, tid = threadIdx.x, N = P.nblk, cta = blockIdx.x, NP = N / 2;
  const int kp = (k + 1) & ~1;                     // padded column length (16-byte aligned columns)
  double* Xs = reinterpret_cast<double*>(smem_raw);             // [NC][kp]   slot s = columns s*BC .. s*BC+BC-1
  double* Js = Xs + (size_t)(WJ ? NC : 0) * kp;                  // [NC][kp] (absent without accumulation)
  double* red = Js + (size_t)NC * kp;                            // [2][BC][RBS]
  int* perm = reinterpret_cast<int*>(red + (size_t)2 * BC * RBS);   // [N] block id at each position
  __shared__ int s_rot, s_big;
  const int team = tid / TS, e = tid % TS;
  const double tol2 = P.tol * P.tol;
  const double thr2 = P.tol / (4.0 * k);           // "tiny rotation": |cos| <= sqrt(tol / 4k)
  for (int i = tid; i < N; i += BC * TS) perm[i] = i;
  int spos[2] = {-1, -1};                          // position held by each slot (-1: slot empty)
  unsigned epoch = 0, gstep = 0;
  int sweep = 0, converged = 0;
  long long tph[4] = {0, 0, 0, 0}, tlast = clock64();
#define JTICK(i) if (tid == 0) { long long _t = clock64(); tph[i] += _t - tlast; tlast = _t; }
  // copy one block between its home in global memory and a slot: ONE thread issues a bulk copy per column (the copy
  // engine moves the 4 KB columns; completion through an mbarrier for loads, a bulk group for stores)
  __shared__ __align__(8) uint64_t s_mbar;
  uint32_t mphase = 0;
  if (tid == 0) {
    jmbar_init(&s_mbar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  const uint32_t colbytes = (uint32_t)kp * 8;
  auto load_slot_bulk = [&](int slot, int blk) {        // tid 0 only; returns the bytes it asked for
    uint32_t bytes = 0;
    for (int c = 0; c < BC; ++c) {
      const int gc = blk * BC + c;
      if (gc < k) {
        bulk_g2s(Xs + (size_t)(slot * BC + c) * kp, P.X + (int64_t)gc * P.ldx, colbytes, &s_mbar);
        if (WJ) bulk_g2s(Js + (size_t)(slot * BC + c) * kp, P.J + (int64_t)gc * P.ldj, colbytes, &s_mbar);
        bytes += (WJ ? 2 : 1) * colbytes;
      }
    }
    return bytes;
  };
  auto zero_slot_tail = [&](int slot, int blk) {        // all threads: columns beyond k are zero
    const int gc = blk * BC + team;
    if (gc >= k) {
      double* xs = Xs + (size_t)(slot * BC + team) * kp;
      double* js = Js + (size_t)(slot * BC + team) * kp;
      for (int r = e; r < kp; r += TS) {
        xs[r] = 0.0;
        if (WJ) js[r] = 0.0;
      }
    }
  };
  auto store_slot_bulk = [&](int slot, int blk) {       // tid 0 only
    for (int c = 0; c < BC; ++c) {
      const int gc = blk * BC + c;
      if (gc < k) {
        bulk_s2g(P.X + (int64_t)gc * P.ldx, Xs + (size_t)(slot * BC + c) * kp, colbytes);
        if (WJ) bulk_s2g(P.J + (int64_t)gc * P.ldj, Js + (size_t)(slot * BC + c) * kp, colbytes);
      }
    }
  };
  // hand-over through the mailbox of a block: [BC][kp] doubles of X (then of J) as LL32 words, two doubles per word
  constexpr int NT = BC * TS;
  const int nwx = BC * kp / 2;                          // words of the X part
  auto send_ll = [&](int slot, int blk, unsigned ver) {
    LL32* mb = P.mbox + (size_t)blk * P.mbox_words;
    const double2* xs = reinterpret_cast<const double2*>(Xs + (size_t)slot * BC * kp);
    for (int w = tid; w < nwx; w += NT) {
      const double2 v = xs[w];
      ll32_store2(mb + w, v.x, v.y, ver);
    }
    if (WJ) {
      const double2* js = reinterpret_cast<const double2*>(Js + (size_t)slot * BC * kp);
      for (int w = tid; w < nwx; w += NT) {
        const double2 v = js[w];
        ll32_store2(mb + nwx + w, v.x, v.y, ver);
      }
    }
  };
  auto fetch_ll = [&](int slot, int blk, unsigned ver) {
    const LL32* mb = P.mbox + (size_t)blk * P.mbox_words;
    constexpr int UB = 8;                               // words in flight per thread
    for (int part = 0; part < (WJ ? 2 : 1); ++part) {
      double2* dst = reinterpret_cast<double2*>((part ? Js : Xs) + (size_t)slot * BC * kp);
      const LL32* src = mb + (part ? nwx : 0);
      for (int w0 = tid; w0 < nwx; w0 += NT * UB) {
        uint32_t q[UB][8];
        unsigned pend = 0;
#pragma unroll
        for (int u = 0; u < UB; ++u)
          if (w0 + u * NT < nwx) pend |= 1u << u;
        // poll in batches: every round re-issues ALL words still pending (one L2 round trip per round, not per word)
        uint32_t spins = 0;
        while (pend) {
#pragma unroll
          for (int u = 0; u < UB; ++u)
            if (pend & (1u << u)) ll32_ld(src + w0 + u * NT, q[u]);
#pragma unroll
          for (int u = 0; u < UB; ++u)
            if ((pend & (1u << u)) && ((q[u][1] ^ ver) | (q[u][3] ^ ver) | (q[u][5] ^ ver) | (q[u][7] ^ ver)) == 0) {
              pend &= ~(1u << u);
              dst[w0 + u * NT] = make_double2(__hiloint2double((int)q[u][2], (int)q[u][0]),
                                              __hiloint2double((int)q[u][6], (int)q[u][4]));
            }
          if (++spins > (1u << 22)) break;              // never hang the box: garbage then fails the convergence check
        }
      }
    }
  };
  auto pair_of = [&](unsigned g, int& pL, int& pR) {       // positions CTA `cta` works on at step g (pL < 0: idle)
    if ((g & 1) == 0) {
      pL = 2 * cta;
      pR = 2 * cta + 1;
    } else if (cta < NP - 1) {
      pL = 2 * cta + 1;
      pR = 2 * cta + 2;
    } else {
      pL = pR = -1;
    }
  };
  __syncthreads();
  for (; sweep < P.max_sweeps; ++sweep) {
    if (tid == 0) {
      s_rot = 0;
      s_big = 0;
    }
    bool rot_any = false;
    for (int st = 0; st < N; ++st, ++gstep) {
      int pL, pR;
      pair_of(gstep, pL, pR);
      if (pL >= 0) {
        // ---- fetch the block(s) this step needs and the CTA does not hold ----
        int sl = (spos[0] == pL) ? 0 : (spos[1] == pL ? 1 : -1);
        int sr = (spos[0] == pR) ? 0 : (spos[1] == pR ? 1 : -1);
        const bool needL = sl < 0, needR = sr < 0;
        if (needL) sl = (sr == 0) ? 1 : 0;
        if (needR) sr = 1 - sl;
        if (needL || needR) {
          if (gstep == 0) {
            // first step of the factorization: the blocks come from their home in global memory (bulk copies)
            JTICK(0)
            if (tid == 0) {
              // expect_tx first (the barrier cannot complete before the byte count is known), then the copies
              uint32_t want = 0;
              if (needL)
                for (int c = 0; c < BC; ++c) want += (perm[pL] * BC + c < k) ? (WJ ? 2 : 1) * colbytes : 0;
              if (needR)
                for (int c = 0; c < BC; ++c) want += (perm[pR] * BC + c < k) ? (WJ ? 2 : 1) * colbytes : 0;
              jmbar_expect_tx(&s_mbar, want);
              if (needL) load_slot_bulk(sl, perm[pL]);
              if (needR) load_slot_bulk(sr, perm[pR]);
            }
            if (needL) zero_slot_tail(sl, perm[pL]);
            if (needR) zero_slot_tail(sr, perm[pR]);
            jmbar_wait(&s_mbar, mphase);
            mphase ^= 1;
            __syncthreads();
            JTICK(1)
          } else {
            // every later step: the neighbour left the block in its mailbox as self-validating words stamped with this
            // step -- ONE L2 hop (poll the data itself), no flag, no second round trip
            if (needL) fetch_ll(sl, perm[pL], gstep);
            if (needR) fetch_ll(sr, perm[pR], gstep);
            __syncthreads();
            JTICK(1)
          }
        }
        spos[sl] = pL;
        spos[sr] = pR;
        int par = 0;
        // ---- once per sweep: the pairs INSIDE each of the two blocks (round-robin over BC columns, both blocks at
        // once: teams 0..BC/2-1 take slot 0, the rest slot 1) ----
        if (st == 0 && BC > 1) {
          for (int rd = 0; rd < BC - 1; ++rd, par ^= 1) {
            int p, q;
            rr_pair(BC, rd, team % (BC / 2 > 0 ? BC / 2 : 1), p, q);
            const int off = (team < BC / 2) ? 0 : BC;
            rot_any |= rotate_pair<TS, R, WJ>(Xs, Js, kp, k, off + p, off + q, team, e, red + (size_t)(par * BC + team) * RBS,
                                          tol2, thr2, &s_big);
            __syncthreads();
          }
        }
        // ---- every step: each column of one block meets each column of the other once (BC rounds of BC disjoint pairs) ----
        for (int rd = 0; rd < BC; ++rd, par ^= 1) {
          const int p = team, q = BC + ((team + rd) % BC);
          rot_any |= rotate_pair<TS, R, WJ>(Xs, Js, kp, k, p, q, team, e, red + (size_t)(par * BC + team) * RBS, tol2, thr2,
                                        &s_big);
          __syncthreads();
        }
        JTICK(2)
        // the two blocks swap positions
        spos[sl] = pR;
        spos[sr] = pL;
      }
      // every CTA replays the whole permutation (N <= a few hundred entries)
      {
        const int npairs = ((gstep & 1) == 0) ? NP : NP - 1, o = (int)(gstep & 1);
        for (int i = tid; i < npairs; i += BC * TS) {
          const int t0 = perm[2 * i + o];
          perm[2 * i + o] = perm[2 * i + o + 1];
          perm[2 * i + o + 1] = t0;
        }
      }
      __syncthreads();
      // ---- hand over: a held block that this CTA does not work on at the next step goes back to global memory ----
      {
        int nL, nR;
        pair_of(gstep + 1, nL, nR);
        bool sent[2] = {false, false};
        int sblk[2] = {0, 0};
        unsigned sver[2] = {0, 0};
#pragma unroll
        for (int sidx = 0; sidx < 2; ++sidx) {
          const int pp = spos[sidx];
          if (pp >= 0 && pp != nL && pp != nR) {
            sent[sidx] = true;
            sblk[sidx] = perm[pp];
            // an end position idles during the next (odd) step: its next reader is two steps away
            const bool idles = (((gstep + 1) & 1) != 0) && (pp == 0 || pp == N - 1);
            sver[sidx] = gstep + 1 + (idles ? 1u : 0u);
            spos[sidx] = -1;
          }
        }
        if (sent[0] || sent[1]) {
#pragma unroll
          for (int sidx = 0; sidx < 2; ++sidx)
            if (sent[sidx]) send_ll(sidx, sblk[sidx], sver[sidx]);
          __syncthreads();       // the slots may be refilled only after they have been read
        }
      }
      JTICK(3)
    }
    // ---- convergence vote: one grid barrier per sweep.  Converged when no pair rotated, or when every rotation of
    // the sweep was tiny (|cos| <= sqrt(tol/4k)): the perturbation a pair receives from the <= 2k later tiny rotations
    // of its two columns is <= 2k * tol/4k = tol/2, so all cosines end below 1.5 tol and the confirming sweep is moot.
    if (rot_any) s_rot = 1;
    __syncthreads();
    if (tid == 0 && s_rot) atomicExch(P.rotated + sweep, 1);
    if (tid == 0 && s_big) atomicExch(P.rotated + P.max_sweeps + sweep, 1);
    ++epoch;
    grid_barrier(P.bar, epoch * gridDim.x);
    int any, big;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(any) : "l"(P.rotated + sweep) : "memory");
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(big) : "l"(P.rotated + P.max_sweeps + sweep) : "memory");
    if (!any || !big) {
      converged = 1;
      ++sweep;
      break;
    }
  }
  // ---- drain: a sweep ends on an odd step, so the next step would be an even one in which CTA i works on positions
  // 2i and 2i+1 -- every block is wanted by exactly one CTA.  Blocks this CTA does not hold sit in their mailboxes
  // (stamp gstep); fetch them, then every block goes home. ----
  if (gstep > 0) {
    int pL, pR;
    pair_of(gstep, pL, pR);
    int sl = (spos[0] == pL) ? 0 : (spos[1] == pL ? 1 : -1);
    int sr = (spos[0] == pR) ? 0 : (spos[1] == pR ? 1 : -1);
    const bool needL = sl < 0, needR = sr < 0;
    if (needL) sl = (sr == 0) ? 1 : 0;
    if (needR) sr = 1 - sl;
    if (needL) fetch_ll(sl, perm[pL], gstep);
    if (needR) fetch_ll(sr, perm[pR], gstep);
    spos[sl] = pL;
    spos[sr] = pR;
  }
  // ---- blocks go home ----
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  if (tid == 0) {
#pragma unroll
    for (int sidx = 0; sidx < 2; ++sidx)
      if (spos[sidx] >= 0) store_slot_bulk(sidx, perm[spos[sidx]]);
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
  // report from a middle CTA (the end CTAs idle or hold the zero-padded block and are not representative)
  if (cta == gridDim.x / 2 && tid == 0) {
    P.out[0] = sweep;
    P.out[1] = converged;
    for (int i = 0; i < 4; ++i) P.out[2 + i] = (int)(tph[i] >> 10);     // kilo-cycles: wait, load, rotate, hand-over
    for (int i = 4; i < 8; ++i) P.out[2 + i] = 0;
  }
#undef JTICK
}

template <int BC, int TS, int R, bool WJ = true>
cudaError_t launch_jacobi(const JacobiParams& P, int grid, size_t smem, cudaStream_t st) {
  cudaError_t e = cudaFuncSetAttribute(jacobi_team_kernel<BC, TS, R, WJ>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  void* args[] = {(void*)&P};
  return cudaLaunchCooperativeKernel((void*)jacobi_team_kernel<BC, TS, R, WJ>, dim3(grid), dim3(BC * TS), args, smem, st);
}

__global__ void set_identity_kernel(int k, double* __restrict__ J, int64_t ldj) {
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < (int64_t)k * k; e += (int64_t)gridDim.x * blockDim.x)
    J[(e % k) + (e / k) * ldj] = ((e % k) == (e / k)) ? 1.0 : 0.0;
}

__global__ void col_norms_kernel(int k, const double* __restrict__ X, int64_t ldx, double* __restrict__ nrm) {
  const int j = blockIdx.x;
  double a = 0.0;
  for (int r = threadIdx.x; r < k; r += 128) a = fma(X[r + (int64_t)j * ldx], X[r + (int64_t)j * ldx], a);
  __shared__ double red[4];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = a;
  __syncthreads();
  if (threadIdx.x == 0) nrm[j] = sqrt(red[0] + red[1] + red[2] + red[3]);
}

}  // namespace

int bra_gather_cols(bra_ctx* ctx, char trans, const double* A, int64_t lda, int64_t mC, int64_t k, const int64_t* idx1,
                    double* C, int64_t ldc) {
  if (k <= 0 || mC <= 0) return BRA_OK;
  ProfScope ps(ctx, BRA_PROF_TAIL);
  gather_cols_kernel<<<(unsigned)std::min<int64_t>(k, 148 * 8), 256, 0, ctx->stream>>>(trans, A, lda, mC, k, idx1, C, ldc);
  ctx->launches++;
  BRA_CUDA(cudaGetLastError());
  return BRA_OK;
}

int bra_transpose(bra_ctx* ctx, const double* src, int64_t lds, int64_t rows, int64_t cols, double* dst, int64_t ldd) {
  if (rows <= 0 || cols <= 0) return BRA_OK;
  ProfScope ps(ctx, BRA_PROF_TAIL);
  dim3 grid((unsigned)((rows + 31) / 32), (unsigned)((cols + 31) / 32));
  transpose2_kernel<<<grid, dim3(32, 8), 0, ctx->stream>>>(src, lds, rows, cols, dst, ldd);
  ctx->launches++;
  BRA_CUDA(cudaGetLastError());
  return BRA_OK;
}

// Y <- Y R^{-1}.  Tall Y: R^{-1} is formed once (k x k back substitution on the identity; for the graded R of a
// pivoted QR the componentwise condition |R^{-1}||R| does not see the grading, so this is as backward stable as
// the substitution) and applied as a GEMM on the TMA + DMMA kernel: (Y R^{-1})' = (R^{-1})' Y'.
int bra_trsolve_right_upper(bra_ctx* ctx, int64_t rows, int k, const double* R, int64_t ldr, double* Y, int64_t ldy) {
  if (rows <= 0 || k <= 0) return BRA_OK;
  const int64_t ldk = (k + 1) & ~int64_t(1);
  if (rows < 4 * (int64_t)k || !bra_gemm_tma_ok(Y, ldy, rows, k)) {
    ProfScope ps(ctx, BRA_PROF_QR);
    trsolve_right_upper_kernel<<<(unsigned)((rows + 63) / 64), 256, 0, ctx->stream>>>(rows, k, R, ldr, Y, ldy);
    ctx->launches++;
    BRA_CUDA(cudaGetLastError());
    return BRA_OK;
  }
  BRA_CUDA(ctx->rinv.reserve((size_t)ldk * k * 8));
  BRA_CUDA(ctx->yt.reserve((size_t)2 * ldk * rows * 8));
  double* Rinv = ctx->rinv.as<double>();
  double* Yt = ctx->yt.as<double>();
  double* Ct = Yt + (size_t)ldk * rows;
  int rc;
  {
    ProfScope ps(ctx, BRA_PROF_QR);
    set_identity_kernel<<<ctx->num_sms * 2, 256, 0, ctx->stream>>>(k, Rinv, ldk);
    ctx->launches++;
    if ((rc = bra_tri_inverse_upper(ctx, k, R, ldr, Rinv, ldk))) return rc;               // Rinv = R^{-1}
  }
  if ((rc = bra_transpose(ctx, Y, ldy, rows, k, Yt, ldk))) return rc;                       // Y' (k x rows)
  if ((rc = bra_gemm_tn(ctx, Rinv, ldk, k, k, Yt, ldk, rows, Ct, ldk))) return rc;          // (Y Rinv)' = Rinv' Y'
  return bra_transpose(ctx, Ct, ldk, k, rows, Y, ldy);
}

// G (k x k, symmetric positive definite, destroyed) -> Rout upper triangular with G = Rout^T Rout
int bra_cholesky_upper(bra_ctx* ctx, int k, double* G, int64_t ldg, double* Rout, int64_t ldr) {
  if (k <= 0) return BRA_OK;
  ProfScope ps(ctx, BRA_PROF_QR);
  // sticky status: reset by bra_chol_status_reset, read by bra_chol_status; the side lane reports in the next word
  int* info = ctx->info.as<int>() + 12 + (ctx->lane ? 1 : 0);
  BRA_CUDA(ctx->ws_cholscr().reserve((size_t)TB * TB * 8 + 256));
  double* Dinv = ctx->ws_cholscr().as<double>();
  unsigned* bar = reinterpret_cast<unsigned*>(Dinv + TB * TB);
  BRA_CUDA(cudaMemset2DAsync(Rout, (size_t)ldr * 8, 0, (size_t)k * 8, (size_t)k, ctx->stream));    // zeros below the diagonal
  BRA_CUDA(cudaMemsetAsync(bar, 0, 4, ctx->stream));
  const int nblk = (k + TB - 1) / TB;
  const int nt0 = (nblk - 1) * nblk / 2;
  int grid = nt0 < 1 ? 1 : nt0;
  // the side lane shares the machine with the main chain: half the SMs each keeps both cooperative grids co-resident
  const int cap = ctx->num_sms / 2 > 0 ? ctx->num_sms / 2 : 1;
  if (grid > cap) grid = cap;
  void* args[] = {(void*)&k, (void*)&G, (void*)&ldg, (void*)&Rout, (void*)&ldr, (void*)&Dinv, (void*)&info, (void*)&bar};
  BRA_CUDA(cudaLaunchCooperativeKernel((void*)chol_fused_kernel, dim3(grid), dim3(256), args, 0, ctx->stream));
  ctx->launches++;
  BRA_CUDA(cudaGetLastError());
  return BRA_OK;
}

// Two Cholesky-QR passes on Y (rows x k, well conditioned): on return Y holds Q (orthonormal columns) and
// Rout (k x k, ld k) = R_y2 * R_y1 * Rpre (Rpre = the preconditioner already divided out of Y, or null).
// defer_last_solve: the second pass stops after its Cholesky -- Y still holds the once-orthogonalised Y1 and the caller
// applies R_y2^{-1} (left in ctx->scratch, k x k, ld k) when and where it wants (psvdfact: on the side lane, next to the
// Jacobi SVD, which only needs Rout).
int bra_cholqr2(bra_ctx* ctx, int64_t rows, int k, double* Y, int64_t ldy, const double* Rpre, double* Rout,
                bool rows_sharded, bool defer_last_solve) {
  if (k <= 0) return BRA_OK;
  BRA_CUDA(ctx->G.reserve((size_t)k * k * 8));
  BRA_CUDA(ctx->scratch.reserve((size_t)3 * k * k * 8));
  double* G = ctx->G.as<double>();
  double* Ry = ctx->scratch.as<double>();          // current pass factor
  double* Racc = Ry + (size_t)k * k;                // accumulated product
  double* Rtmp = Racc + (size_t)k * k;
  int rc;
  for (int pass = 0; pass < 2; ++pass) {
    if ((rc = bra_gemm_tn(ctx, Y, ldy, k, rows, Y, ldy, k, G, k))) return rc;          // G = Y^T Y
    if (rows_sharded && (rc = bra_allreduce_sum_f64(ctx, G, (int64_t)k * k))) return rc;   // sum over the row blocks
    if ((rc = bra_cholesky_upper(ctx, k, G, k, Ry, k))) return rc;
    if (!(defer_last_solve && pass == 1) && (rc = bra_trsolve_right_upper(ctx, rows, k, Ry, k, Y, ldy))) return rc;   // Y <- Y Ry^{-1}
    // Racc <- Ry * (pass == 0 ? Rpre : Racc)     (k x k upper-triangular products)
    const double* prev = (pass == 0) ? Rpre : Racc;
    if (prev) {
      // C = Ry * prev : generic strided GEMM, Om(i,kk) = Ry[i + kk*k], Aop(kk,j) = prev[kk + j*k]
      if ((rc = bra_gemm_generic(ctx, Ry, 1, k, prev, 1, k, k, k, k, Rtmp, k))) return rc;
      BRA_CUDA(cudaMemcpyAsync(Racc, Rtmp, (size_t)k * k * 8, cudaMemcpyDeviceToDevice, ctx->stream));
    } else {
      BRA_CUDA(cudaMemcpyAsync(Racc, Ry, (size_t)k * k * 8, cudaMemcpyDeviceToDevice, ctx->stream));
    }
  }
  BRA_CUDA(cudaMemcpyAsync(Rout, Racc, (size_t)k * k * 8, cudaMemcpyDeviceToDevice, ctx->stream));
  return BRA_OK;      // a non-positive pivot is sticky in ctx->info[12]: the caller checks it at its next sync
}

int bra_chol_status_reset(bra_ctx* ctx) {
  BRA_CUDA(ctx->info.reserve(64));
  BRA_CUDA(cudaMemsetAsync(ctx->info.as<int>() + 12 + (ctx->lane ? 1 : 0), 0, 4, ctx->stream));
  return BRA_OK;
}

// ---- side lane: launches between bra_lane_fork and bra_lane_end go to a second stream that starts after everything
// queued on the main stream so far; bra_lane_join makes the main stream wait for them ----
int bra_lane_fork(bra_ctx* ctx) {
  if (!ctx->side_stream) {
    BRA_CUDA(cudaStreamCreateWithFlags(&ctx->side_stream, cudaStreamNonBlocking));
    BRA_CUDA(cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming));
    BRA_CUDA(cudaEventCreateWithFlags(&ctx->ev_join, cudaEventDisableTiming));
  }
  BRA_CUDA(cudaEventRecord(ctx->ev_fork, ctx->stream));
  BRA_CUDA(cudaStreamWaitEvent(ctx->side_stream, ctx->ev_fork, 0));
  ctx->lane_saved = ctx->stream;
  ctx->stream = ctx->side_stream;
  ctx->lane = 1;
  return BRA_OK;
}
int bra_lane_end(bra_ctx* ctx) {
  if (!ctx->lane) return BRA_OK;
  cudaError_t e = cudaEventRecord(ctx->ev_join, ctx->side_stream);
  ctx->stream = ctx->lane_saved;
  ctx->lane = 0;
  BRA_CUDA(e);
  return BRA_OK;
}
int bra_lane_join(bra_ctx* ctx) {
  BRA_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->ev_join, 0));
  return BRA_OK;
}

// Synchronises the stream and reports a Cholesky breakdown recorded since the last reset.
int bra_chol_status(bra_ctx* ctx) {
  BRA_CUDA(cudaMemcpyAsync(ctx->h_info + 12, ctx->info.as<int>() + 12, 4, cudaMemcpyDeviceToHost, ctx->stream));
  BRA_CUDA(cudaStreamSynchronize(ctx->stream));
  if (ctx->h_info[12] != 0) {
    ctx->set_error("CholeskyQR2: Gram matrix not positive definite at pivot " + std::to_string(ctx->h_info[12]));
    return BRA_ERR_INTERNAL;
  }
  return BRA_OK;
}

// One-sided Jacobi on the columns of X (k x k): X J = U diag(sigma).  On return X holds U diag(sigma) (columns
// unsorted), J the accumulated rotations, sigma_host the column norms, order_host the descending order.
// skip_J (optional, in/out): in = the caller can do without the accumulated rotations (it recovers them from a
// triangular X); out = whether the kernel really ran without J (only the 8-rows-per-thread configurations do).
int bra_jacobi_svd(bra_ctx* ctx, int k, double* X, int64_t ldx, double* J, int64_t ldj, double* sigma_host,
                   int* order_host, bool* skip_J) {
  const bool may_skip = skip_J && *skip_J && getenv("BRA_JACOBI_KEEPJ") == nullptr;
  if (skip_J) *skip_J = false;
  if (k <= 0) return BRA_OK;
  ProfScope ps(ctx, BRA_PROF_SVD);
  constexpr int MAX_SWEEPS = 48;
  int sweeps = 0, converged = 1;
  bool have_out = false;
  if ((ldx & 1) || (ldj & 1) || (reinterpret_cast<uintptr_t>(X) & 15) || (reinterpret_cast<uintptr_t>(J) & 15)) {
    ctx->set_error("Jacobi SVD: X and J need even leading dimensions and 16-byte aligned bases");
    return -4;
  }
  if (k == 1) {
    set_identity_kernel<<<1, 32, 0, ctx->stream>>>(k, J, ldj);
    ctx->launches++;
  }
  if (k > 1) {
    // team size: k <= TS * R rows in registers (R = 4, or 8 on request); panel = 2*BC columns of X and of J in
    // shared memory
    int ts = 32, rr = 4, bc = 4;
    bool noj = false;
    const size_t kp = (size_t)((k + 1) & ~1);
    const size_t budget = (size_t)ctx->smem_optin - 2048;
    auto need = [&](int b) { return (size_t)(noj ? 16 : 32) * b * kp + (size_t)2 * b * ((ts / 32) * 4 + 2) * 8 + (size_t)4 * (2 * ((k + 2 * b - 1) / (2 * b))) + 64; };
    int nblk;
    if (k > 1024) {
      // large cores: one CTA per block pair must be co-resident (k <= 2 bc #SMs) and the panel must fit shared memory
      // (2 bc columns of X, and of J unless the caller recovers the rotations itself).  Without J: blocks of 6 columns,
      // 4 warps x 16 rows per pair (k <= 12 #SMs = 1776 on B200); with J: blocks of 4, 5 warps x 8 rows
      // (k <= 8 #SMs = 1184, 1280 rows).
      noj = may_skip;
      if (noj) {
        bc = 6; ts = 128; rr = 16;
      } else {
        bc = 4; ts = 160; rr = 8;
      }
      nblk = 2 * ((k + 2 * bc - 1) / (2 * bc));
      if (nblk / 2 > ctx->num_sms || need(bc) > budget || ts * rr < k) {
        ctx->set_error("Jacobi SVD: core too large -- one CTA per pair of column blocks must be co-resident: k <= 12 * #SMs "
                       "(1776 on B200) for psvdfact, k <= 8 * #SMs (1184) where the rotations are accumulated (pheigfact, CUR)");
        return BRA_ERR_UNSUPPORTED;
      }
      if (skip_J) *skip_J = noj;
      if (!noj) {
        set_identity_kernel<<<ctx->num_sms * 2, 256, 0, ctx->stream>>>(k, J, ldj);
        ctx->launches++;
      }
    } else {
      while (ts < 1024 && ts * 4 < k) ts <<= 1;
      // measured on B200 (k = 500): 8 rows per thread and blocks of 4 columns (63 CTAs) balance the per-round
      // latency chain against the per-step exchange
      bool r8 = ts >= 64 && ts <= 256;
      if (const char* ev = getenv("BRA_JACOBI_R8")) r8 = r8 && atoi(ev) > 0;
      if (r8) {
        ts >>= 1;
        rr = 8;
      }
      noj = may_skip && rr == 8;
      // without the J panel a pair fits one warp (16 rows per lane): no cross-warp reduction, no team barrier --
      // measured at k = 498: 4.37 ms against 4.56 ms with two warps per pair
      const char* r16 = getenv("BRA_JACOBI_R16");
      if (noj && k <= 512 && !(r16 && atoi(r16) == 0)) {
        ts = 32;
        rr = 16;
      }
      if (skip_J) *skip_J = noj;
      if (!noj) {
        set_identity_kernel<<<ctx->num_sms * 2, 256, 0, ctx->stream>>>(k, J, ldj);
        ctx->launches++;
      }
      bc = 1024 / ts;
      if (bc > 4) bc = 4;
      if (const char* ev = getenv("BRA_JACOBI_BC")) {
        const int v = atoi(ev);
        if (v > 0 && v <= 8 && v * ts <= 1024 && (v & (v - 1)) == 0) bc = v;
      }
      while (bc > 1 && (need(bc) > budget || k <= bc)) bc >>= 1;   // tiny cores: keep at least two blocks of real columns
      if (need(bc) > budget) {
        ctx->set_error("Jacobi SVD: k too large for the shared-memory panel");
        return BRA_ERR_UNSUPPORTED;
      }
      nblk = 2 * ((k + 2 * bc - 1) / (2 * bc));
      while (nblk / 2 > ctx->num_sms && bc < 8) {      // one CTA per block pair must be co-resident
        bc <<= 1;
        nblk = 2 * ((k + 2 * bc - 1) / (2 * bc));
      }
      if (nblk / 2 > ctx->num_sms || need(bc) > budget || bc * ts > 1024) {
        ctx->set_error("Jacobi SVD: core does not fit the co-resident grid");
        return BRA_ERR_UNSUPPORTED;
      }
    }
    const int grid = nblk / 2;
    const size_t smem = need(bc);
    const size_t wbytes0 = ((1024 + (size_t)2 * MAX_SWEEPS * 4 + (size_t)nblk * 4) + 255) & ~size_t(255);
    const size_t mbox_words = (size_t)bc * kp / 2 * (noj ? 1 : 2);
    const size_t wbytes = wbytes0 + (size_t)nblk * mbox_words * sizeof(LL32);
    BRA_CUDA(ctx->jwork.reserve(wbytes));
    BRA_CUDA(cudaMemsetAsync(ctx->jwork.p, 0, wbytes, ctx->stream));      // stamps 0 are invalid: the steps count from 1
    JacobiParams P;
    P.k = k;
    P.nblk = nblk;
    P.X = X;
    P.ldx = ldx;
    P.J = J;
    P.ldj = ldj;
    P.tol = std::sqrt((double)k) * 1.1102230246251565e-16;      // sqrt(k) * eps, dgesvj's criterion
    P.max_sweeps = MAX_SWEEPS;
    P.bar = ctx->jwork.as<unsigned>();
    P.out = ctx->jwork.as<int>() + 16;
    P.rotated = ctx->jwork.as<int>() + 64;
    P.bstep = ctx->jwork.as<unsigned>() + 256 + 2 * MAX_SWEEPS;
    P.mbox = reinterpret_cast<LL32*>(reinterpret_cast<unsigned char*>(ctx->jwork.p) + wbytes0);
    P.mbox_words = mbox_words;
    cudaError_t e = cudaErrorInvalidValue;
#define JL(B_, T_, R_) if (bc == B_ && ts == T_ && rr == R_ && !noj) e = launch_jacobi<B_, T_, R_>(P, grid, smem, ctx->stream);
#define JN(B_, T_) if (bc == B_ && ts == T_ && noj && rr == 8) e = launch_jacobi<B_, T_, 8, false>(P, grid, smem, ctx->stream);
    JL(8, 32, 4) JL(4, 32, 4) JL(2, 32, 4) JL(1, 32, 4)
    JL(8, 64, 4) JL(4, 64, 4) JL(2, 64, 4) JL(1, 64, 4)
    JL(8, 128, 4) JL(4, 128, 4) JL(2, 128, 4) JL(1, 128, 4)
    JL(4, 256, 4) JL(2, 256, 4) JL(1, 256, 4)
    JL(2, 512, 4) JL(1, 512, 4)
    JL(1, 1024, 4)
    JL(8, 32, 8) JL(4, 32, 8) JL(8, 64, 8) JL(4, 64, 8) JL(8, 128, 8) JL(4, 128, 8)
    JN(8, 32) JN(4, 32) JN(8, 64) JN(4, 64) JN(8, 128) JN(4, 128)
    if (noj && rr == 16 && bc == 4) e = launch_jacobi<4, 32, 16, false>(P, grid, smem, ctx->stream);
    if (noj && rr == 16 && bc == 8) e = launch_jacobi<8, 32, 16, false>(P, grid, smem, ctx->stream);
    if (noj && rr == 16 && bc == 6 && ts == 128) e = launch_jacobi<6, 128, 16, false>(P, grid, smem, ctx->stream);
    if (!noj && rr == 8 && bc == 4 && ts == 160) e = launch_jacobi<4, 160, 8, true>(P, grid, smem, ctx->stream);
#undef JL
#undef JN
    BRA_CUDA(e);
    ctx->launches++;
    have_out = true;
  }
  // column norms = singular values; ONE host round trip for the kernel's report and the k norms (pinned scratch)
  BRA_CUDA(ctx->S.reserve((size_t)k * 8));
  col_norms_kernel<<<k, 128, 0, ctx->stream>>>(k, X, ldx, ctx->S.as<double>());
  ctx->launches++;
  int* h = reinterpret_cast<int*>(ctx->h_pin);
  double* hs = reinterpret_cast<double*>(ctx->h_pin + 64);
  if ((size_t)k * 8 + 64 > BRA_HPIN_BYTES) {
    ctx->set_error("Jacobi SVD: k too large for the pinned read-back buffer");
    return BRA_ERR_UNSUPPORTED;
  }
  if (have_out) BRA_CUDA(cudaMemcpyAsync(h, ctx->jwork.as<int>() + 16, 40, cudaMemcpyDeviceToHost, ctx->stream));
  BRA_CUDA(cudaMemcpyAsync(hs, ctx->S.p, (size_t)k * 8, cudaMemcpyDeviceToHost, ctx->stream));
  BRA_CUDA(cudaStreamSynchronize(ctx->stream));
  if (have_out) {
    sweeps = h[0];
    converged = h[1];
    for (int i = 0; i < 8; ++i) ctx->jacobi_kcycles[i] = h[2 + i];
  }
  ctx->last_jacobi_sweeps = sweeps;
  std::memcpy(sigma_host, hs, (size_t)k * 8);
  std::iota(order_host, order_host + k, 0);
  std::stable_sort(order_host, order_host + k, [&](int a, int b) { return sigma_host[a] > sigma_host[b]; });
  if (!converged) {
    ctx->set_error("Jacobi SVD did not converge in 48 sweeps");
    return BRA_ERR_INTERNAL;
  }
  return BRA_OK;
}

namespace {
// out[:, j] = X[:, order[j]] * (scale ? 1/scale[order[j]] : 1)
__global__ void gather_scale_cols_kernel(const double* __restrict__ X, int64_t ldx, int64_t rows, int kk,
                                         const int* __restrict__ order, const double* __restrict__ scale,
                                         double* __restrict__ out, int64_t ldo) {
  for (int j = blockIdx.x; j < kk; j += gridDim.x) {
    const int s = order[j];
    const double f = scale ? (scale[s] != 0.0 ? 1.0 / scale[s] : 0.0) : 1.0;
    for (int64_t r = threadIdx.x; r < rows; r += blockDim.x) out[r + (int64_t)j * ldo] = X[r + (int64_t)s * ldx] * f;
  }
}
// dst[:, jpvt[j]-1] = src[:, j]
__global__ void scatter_cols_kernel(const double* __restrict__ src, int64_t lds, int64_t rows, int64_t n,
                                    const int64_t* __restrict__ jpvt1, double* __restrict__ dst, int64_t ldd) {
  for (int64_t j = blockIdx.x; j < n; j += gridDim.x) {
    const double* s = src + j * lds;
    double* d = dst + (jpvt1[j] - 1) * ldd;
    for (int64_t r = threadIdx.x; r < rows; r += blockDim.x) d[r] = s[r];
  }
}
}  // namespace

namespace {
// make diag(R) >= 0: row i of R and column i of Q are flipped together (Q R unchanged)
__global__ void fix_signs_kernel(int64_t rows, int k, double* __restrict__ Q, int64_t ldq, double* __restrict__ R,
                                 int64_t ldr) {
  for (int i = blockIdx.x; i < k; i += gridDim.x) {
    if (!(R[i + (int64_t)i * ldr] < 0.0)) continue;      // uniform per block: every thread reads the same entry
    __syncthreads();
    for (int64_t r = threadIdx.x; r < rows; r += blockDim.x) Q[r + (int64_t)i * ldq] = -Q[r + (int64_t)i * ldq];
    for (int j = i + threadIdx.x; j < k; j += blockDim.x) R[i + (int64_t)j * ldr] = -R[i + (int64_t)j * ldr];
    __syncthreads();
  }
}
}  // namespace

int bra_fix_signs(bra_ctx* ctx, int64_t rows, int k, double* Q, int64_t ldq, double* R, int64_t ldr) {
  if (k <= 0) return BRA_OK;
  fix_signs_kernel<<<std::min(k, 148 * 4), 256, 0, ctx->stream>>>(rows, k, Q, ldq, R, ldr);
  ctx->launches++;
  BRA_CUDA(cudaGetLastError());
  return BRA_OK;
}

int bra_set_identity(bra_ctx* ctx, int k, double* J, int64_t ldj) {
  if (k <= 0) return BRA_OK;
  set_identity_kernel<<<ctx->num_sms * 2, 256, 0, ctx->stream>>>(k, J, ldj);
  ctx->launches++;
  BRA_CUDA(cudaGetLastError());
  return BRA_OK;
}

int bra_gather_scale_cols(bra_ctx* ctx, const double* X, int64_t ldx, int64_t rows, int kk, const int* order_dev,
                          const double* scale_dev, double* out, int64_t ldo) {
  if (kk <= 0 || rows <= 0) return BRA_OK;
  gather_scale_cols_kernel<<<std::min(kk, 148 * 8), 128, 0, ctx->stream>>>(X, ldx, rows, kk, order_dev, scale_dev, out, ldo);
  ctx->launches++;
  BRA_CUDA(cudaGetLastError());
  return BRA_OK;
}

int bra_scatter_cols(bra_ctx* ctx, const double* src, int64_t lds, int64_t rows, int64_t n, const int64_t* jpvt1,
                     double* dst, int64_t ldd) {
  if (n <= 0 || rows <= 0) return BRA_OK;
  scatter_cols_kernel<<<(unsigned)std::min<int64_t>(n, 148 * 8), 128, 0, ctx->stream>>>(src, lds, rows, n, jpvt1, dst, ldd);
  ctx->launches++;
  BRA_CUDA(cudaGetLastError());
  return BRA_OK;
}

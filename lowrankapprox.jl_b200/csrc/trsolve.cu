// ID solve T = R11^{-1} R12  (reference: maxdet_t, src/pqr.jl:438-442 -> BLAS dtrsm L,U,N,N).
//
// Column-panel parallel blocked back substitution: a CTA owns 32 right-hand
// sides and walks the 32-row blocks of R11 bottom-up.  The off-diagonal work is
// 32x32x32 tile products out of shared memory; the diagonal block is solved by
// genuine substitution (not by an explicit inverse: the diagonal blocks of a
// graded R11 can be as ill-conditioned as R11 itself).
#include "common.cuh"

namespace {

constexpr int TB = 32;

// NC = right-hand sides per CTA (32: the ID solve, thousands of columns; 8: narrow problems, where 32-wide panels
// would leave most of the machine idle).
template <int NC>
__global__ void __launch_bounds__(256) trsolve_upper_kernel(int k, int64_t nrhs, const double* __restrict__ R,
                                                            int64_t ldr, double* __restrict__ X, int64_t ldx) {
  constexpr int NG = 256 / NC;        // row groups
  constexpr int RPT = TB / NG;        // rows per thread in a 32-row block
  static_assert(RPT >= 1 && RPT * NG == TB, "NC must be 8, 16 or 32");
  __shared__ double Rs[TB][TB + 1];   // Rs[r][c] = R[ib*32 + r, jb*32 + c]
  __shared__ double Xs[TB][NC + 1];   // Xs[r][c] = X[jb*32 + r, col0 + c]
  __shared__ double Acc[TB][NC + 1];
  const int tid = threadIdx.x;
  const int tx = tid % NC;            // rhs column within the panel
  const int ty = tid / NC;            // row group: rows ty*RPT .. ty*RPT+RPT-1
  const int64_t col0 = (int64_t)blockIdx.x * NC;
  const int nblk = (k + TB - 1) / TB;
  const int ib_top = nblk - 1;

  for (int ib = ib_top; ib >= 0; --ib) {
    const int r0 = ib * TB;
    double acc[RPT];
#pragma unroll
    for (int u = 0; u < RPT; ++u) {
      const int r = r0 + ty * RPT + u;
      acc[u] = (r < k && col0 + tx < nrhs) ? X[r + (col0 + tx) * ldx] : 0.0;
    }
    for (int jb = ib + 1; jb <= ib_top; ++jb) {
      const int c0 = jb * TB;
      __syncthreads();
      for (int e = tid; e < TB * TB; e += 256) {
        const int rr = e & 31, cc = e >> 5;
        const int r = r0 + rr, c = c0 + cc;
        Rs[rr][cc] = (r < k && c < k) ? R[r + (int64_t)c * ldr] : 0.0;
      }
      for (int e = tid; e < TB * NC; e += 256) {
        // X rows of block jb were finalised earlier by this same CTA
        const int rr = e % TB, cc = e / TB;
        const int xr = c0 + rr;
        Xs[rr][cc] = (xr < k && col0 + cc < nrhs) ? X[xr + (col0 + cc) * ldx] : 0.0;
      }
      __syncthreads();
#pragma unroll 8
      for (int kk = 0; kk < TB; ++kk) {
        const double x = Xs[kk][tx];
#pragma unroll
        for (int u = 0; u < RPT; ++u) acc[u] = fma(-Rs[ty * RPT + u][kk], x, acc[u]);
      }
    }
    __syncthreads();
    // diagonal block: load R_ii, park acc in smem, NC threads substitute (thread = rhs column)
    for (int e = tid; e < TB * TB; e += 256) {
      const int rr = e & 31, cc = e >> 5;
      const int r = r0 + rr, c = r0 + cc;
      Rs[rr][cc] = (r < k && c < k) ? R[r + (int64_t)c * ldr] : (rr == cc ? 1.0 : 0.0);
    }
#pragma unroll
    for (int u = 0; u < RPT; ++u) Acc[ty * RPT + u][tx] = acc[u];
    __syncthreads();
    if (ty == 0) {
      double x[TB];
#pragma unroll
      for (int r = 0; r < TB; ++r) x[r] = Acc[r][tx];
#pragma unroll
      for (int r = TB - 1; r >= 0; --r) {
        x[r] = x[r] / Rs[r][r];
#pragma unroll
        for (int rr = 0; rr < r; ++rr) x[rr] = fma(-Rs[rr][r], x[r], x[rr]);
      }
      if (col0 + tx < nrhs) {
#pragma unroll
        for (int r = 0; r < TB; ++r)
          if (r0 + r < k) X[(r0 + r) + (col0 + tx) * ldx] = x[r];
      }
    }
    __syncthreads();
  }
}

}  // namespace

// Wide variant for the ID solve proper (thousands of right-hand sides): 64 columns per CTA (twice the reuse of every
// R tile), the next R / X tiles are loaded while the current product runs (register-staged double buffering), and the
// diagonal block is solved with reciprocals of its diagonal formed once per block (what BLAS trsm kernels do) instead
// of one dependent division per row.  510 -> ~150 us at k = 500, n = 8192.
constexpr int WC = 64;                       // right-hand sides per CTA
constexpr int WS = TB + 1;                   // odd stride of the X tiles: conflict-free 8-byte accesses both ways
constexpr int RS = TB + 2;                   // even stride of the R tiles: their reads are broadcasts, taken as 16-byte pairs

__global__ void __launch_bounds__(256) trsolve_upper_wide_kernel(int k, int64_t nrhs, const double* __restrict__ R,
                                                                 int64_t ldr, double* __restrict__ X, int64_t ldx) {
  extern __shared__ __align__(16) double sm[];
  double* Rt = sm;                           // [2][TB cols][RS]   Rt[c][r] = R[ib*32 + r, jb*32 + c]
  double* Xt = Rt + 2 * TB * RS;             // [2][WC cols][WS]   Xt[c][r] = X[jb*32 + r, col0 + c]
  double* Acc = Xt + 2 * WC * WS;            // [WC][WS]
  double* Rd = Acc + WC * WS;                // [TB] reciprocals of the diagonal
  const int tid = threadIdx.x;
  const int tx = tid % WC, ty = tid / WC;    // column, row group (rows ty*8 .. ty*8+7)
  const int64_t col0 = (int64_t)blockIdx.x * WC;
  const int nblk = (k + TB - 1) / TB;

  // register-staged double buffering: the loads of tile jb+1 are issued before the product of tile jb and parked in
  // shared memory after it (8-byte cp.async tops out near 24 B/clk/SM on this part; plain coalesced loads do not)
  double pr[4], px[8];
  auto fetch = [&](int ib, int jb) {
    const int r0 = ib * TB, c0 = jb * TB;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int e = tid + 256 * i, rr = e & 31, cc = e >> 5;
      pr[i] = (r0 + rr < k && c0 + cc < k) ? R[(r0 + rr) + (int64_t)(c0 + cc) * ldr] : 0.0;
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int e = tid + 256 * i, rr = e & 31, cc = e >> 5;
      px[i] = (c0 + rr < k && col0 + cc < nrhs) ? X[(c0 + rr) + (col0 + cc) * ldx] : 0.0;
    }
  };
  auto park = [&](int buf) {
    double* rt = Rt + buf * TB * RS;
    double* xt = Xt + buf * WC * WS;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int e = tid + 256 * i;
      rt[(e >> 5) * RS + (e & 31)] = pr[i];
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int e = tid + 256 * i;
      xt[(e >> 5) * WS + (e & 31)] = px[i];
    }
  };

  for (int ib = nblk - 1; ib >= 0; --ib) {
    const int r0 = ib * TB;
    double acc[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int r = r0 + ty * 8 + u;
      acc[u] = (r < k && col0 + tx < nrhs) ? X[r + (col0 + tx) * ldx] : 0.0;
    }
    // rows of X in blocks jb > ib were finalised earlier by this same CTA (visible after the barrier below)
    __syncthreads();
    if (ib + 1 < nblk) {
      fetch(ib, ib + 1);
      park(0);
    }
    for (int jb = ib + 1; jb < nblk; ++jb) {
      const int buf = (jb - ib - 1) & 1;
      __syncthreads();                                   // tile jb parked; everyone is done with the other buffer
      if (jb + 1 < nblk) fetch(ib, jb + 1);
      const double* rt = Rt + buf * TB * RS + ty * 8;
      const double* xt = Xt + buf * WC * WS + tx * WS;
#pragma unroll 8
      for (int kk = 0; kk < TB; ++kk) {
        const double x = xt[kk];
        const double2* r2 = reinterpret_cast<const double2*>(rt + kk * RS);       // 8 rows of R as four 16-byte loads
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const double2 rv = r2[u];
          acc[2 * u] = fma(-rv.x, x, acc[2 * u]);
          acc[2 * u + 1] = fma(-rv.y, x, acc[2 * u + 1]);
        }
      }
      if (jb + 1 < nblk) park(buf ^ 1);
    }
    __syncthreads();
    // diagonal block: R_ii into Rt[0] (as Rt[c][r]), reciprocals of its diagonal, acc parked in Acc
    for (int e = tid; e < TB * TB; e += 256) {
      const int rr = e & 31, cc = e >> 5;
      const int r = r0 + rr, c = r0 + cc;
      Rt[cc * RS + rr] = (r < k && c < k) ? R[r + (int64_t)c * ldr] : (rr == cc ? 1.0 : 0.0);
    }
    if (tid < TB) Rd[tid] = (r0 + tid < k) ? 1.0 / R[(r0 + tid) + (int64_t)(r0 + tid) * ldr] : 1.0;
#pragma unroll
    for (int u = 0; u < 8; ++u) Acc[tx * WS + ty * 8 + u] = acc[u];
    __syncthreads();
    if (ty == 0) {
      double x[TB];
#pragma unroll
      for (int r = 0; r < TB; ++r) x[r] = Acc[tx * WS + r];
#pragma unroll
      for (int r = TB - 1; r >= 0; --r) {
        x[r] *= Rd[r];
#pragma unroll
        for (int rr = 0; rr < r; ++rr) x[rr] = fma(-Rt[r * RS + rr], x[r], x[rr]);
      }
      if (col0 + tx < nrhs) {
#pragma unroll
        for (int r = 0; r < TB; ++r)
          if (r0 + r < k) X[(r0 + r) + (col0 + tx) * ldx] = x[r];
      }
    }
    __syncthreads();
  }
}

// In place: X (k x nrhs, holds R12 on entry) <- R11^{-1} X
int bra_trsolve_upper(bra_ctx* ctx, int k, int64_t nrhs, const double* R11, int64_t ldr, double* X, int64_t ldx) {
  if (k <= 0 || nrhs <= 0) return BRA_OK;
  if (nrhs >= 32 * (int64_t)ctx->num_sms) {
    const size_t smem = ((size_t)2 * TB * RS + 2 * WC * WS + WC * WS + TB) * 8;
    BRA_CUDA(cudaFuncSetAttribute(trsolve_upper_wide_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    trsolve_upper_wide_kernel<<<(unsigned)((nrhs + WC - 1) / WC), 256, smem, ctx->stream>>>(k, nrhs, R11, ldr, X, ldx);
  } else {
    trsolve_upper_kernel<8><<<(unsigned)((nrhs + 7) / 8), 256, 0, ctx->stream>>>(k, nrhs, R11, ldr, X, ldx);
  }
  ctx->launches++;
  BRA_CUDA(cudaGetLastError());
  return BRA_OK;
}

// ---- blocked recursive inverse of an upper-triangular matrix --------------------------------------------------------
// R = [A B; 0 C]  =>  R^{-1} = [A^{-1}  -A^{-1} B C^{-1}; 0  C^{-1}].  Level 0 inverts the 32 x 32 diagonal blocks (one
// warp each, column-oriented substitution in registers); level s = 32, 64, ... joins neighbouring s-blocks with two
// tiled products (tmp = B C^{-1}, V12 = -A^{-1} tmp), one CTA per 32 x 32 output tile.  log2(k/32) levels of fully
// parallel work instead of every column panel walking all block rows one after another (250 us -> ~60 us at k = 500).
namespace {

struct TriinvSmem {
  double Xs[TB][TB + 1];
  double Ys[TB][TB + 1];
};
// one warp; `U` aliases the CTA's tile buffer
__device__ __forceinline__ void triinv_diag_body(double (*U)[TB + 1], int vb, int k, const double* __restrict__ R, int64_t ldr,
                                                 double* V, int64_t ldv) {
  const int lane = threadIdx.x & 31, b0 = vb * TB;
  const int bs = min(TB, k - b0);
  for (int c = 0; c < TB; ++c)
    U[lane][c] = (lane < bs && c < bs) ? R[(b0 + lane) + (int64_t)(b0 + c) * ldr] : (lane == c ? 1.0 : 0.0);
  __syncwarp();
  const double rd = 1.0 / U[lane][lane];          // lane c keeps 1 / U[c][c]
  // lane j solves U v = e_j: rhs in registers, column-oriented (independent FMAs per eliminated unknown)
  double v[TB];
#pragma unroll
  for (int r = 0; r < TB; ++r) v[r] = (r == lane) ? 1.0 : 0.0;
#pragma unroll
  for (int c = TB - 1; c >= 0; --c) {
    v[c] *= __shfl_sync(0xffffffffu, rd, c);
#pragma unroll
    for (int r = 0; r < c; ++r) v[r] = fma(-U[r][c], v[c], v[r]);
  }
  if (lane < bs) {
#pragma unroll
    for (int r = 0; r < TB; ++r)
      if (r < bs) V[(b0 + r) + (int64_t)(b0 + lane) * ldv] = v[r];
  }
}

// phase 0: tmp[a0+i, a0+s+j] = sum_t R[a0+i, a0+s+t] * V[a0+s+t, a0+s+j]      (t <= j: V upper triangular)
// phase 1: V[a0+i, a0+s+j]   = - sum_t V[a0+i, a0+t] * tmp[a0+t, a0+s+j]      (t >= i)
// (V and tmp are written by other CTAs in earlier phases of the fused kernel: L2 loads)
__device__ __forceinline__ void triinv_join_body(TriinvSmem& sm, int tile, int k, int s, int phase, const double* __restrict__ R,
                                                 int64_t ldr, double* V, int64_t ldv, double* tmp, int64_t ldt) {
  auto& Xs = sm.Xs;
  auto& Ys = sm.Ys;
  const int nt = s / TB;                               // tiles per side of an s-block
  const int pair = tile / (nt * nt), tt = tile % (nt * nt);
  const int ti = tt % nt, tj = tt / nt;
  const int a0 = pair * 2 * s;
  const int r0 = a0 + ti * TB, c0 = a0 + s + tj * TB;  // top-left of the output tile
  if (r0 >= k || c0 >= k) return;
  const int tid = threadIdx.x, tx = tid & 31, ty = tid >> 5;
  double acc[4] = {0.0, 0.0, 0.0, 0.0};
  const int t_lo = (phase == 0) ? 0 : ti, t_hi = (phase == 0) ? tj : nt - 1;
  for (int t = t_lo; t <= t_hi; ++t) {
    // X tile: rows r0.., cols xk0..;  Y tile: rows yk0.., cols c0..
    const double* Xm = (phase == 0) ? R : V;
    const int64_t ldxm = (phase == 0) ? ldr : ldv;
    const int xk0 = (phase == 0) ? a0 + s + t * TB : a0 + t * TB;
    const double* Ym = (phase == 0) ? V : tmp;
    const int64_t ldym = (phase == 0) ? ldv : ldt;
    const int yk0 = xk0;
    __syncthreads();
    for (int e = tid; e < TB * TB; e += 256) {
      const int rr = e & 31, cc = e >> 5;
      Xs[rr][cc] = (r0 + rr < k && xk0 + cc < k) ? __ldcg(Xm + (r0 + rr) + (int64_t)(xk0 + cc) * ldxm) : 0.0;
      Ys[rr][cc] = (yk0 + rr < k && c0 + cc < k) ? __ldcg(Ym + (yk0 + rr) + (int64_t)(c0 + cc) * ldym) : 0.0;
    }
    __syncthreads();
#pragma unroll 8
    for (int kk = 0; kk < TB; ++kk) {
      const double x = Xs[tx][kk];
#pragma unroll
      for (int u = 0; u < 4; ++u) acc[u] = fma(x, Ys[kk][ty * 4 + u], acc[u]);
    }
  }
  double* out = (phase == 0) ? tmp : V;
  const int64_t ldo = (phase == 0) ? ldt : ldv;
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const int r = r0 + tx, c = c0 + ty * 4 + u;
    if (r < k && c < k) out[r + (int64_t)c * ldo] = (phase == 0) ? acc[u] : -acc[u];
  }
}

// the whole blocked inverse in ONE cooperative launch (round 1: one launch for the diagonal blocks + two per level, 9 at
// k = 497): grid barriers between the levels; `bar` is zeroed before the launch
__device__ __forceinline__ bool tri_grid_barrier(unsigned* bar, unsigned target, int* s_fail) {
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

__global__ void __launch_bounds__(256, 1) triinv_fused_kernel(int k, const double* __restrict__ R, int64_t ldr, double* V,
                                                              int64_t ldv, double* tmp, int64_t ldt, unsigned* bar) {
  __shared__ TriinvSmem sm;
  __shared__ int s_fail;
  if (threadIdx.x == 0) s_fail = 0;
  const int nblk = (k + TB - 1) / TB, Gn = gridDim.x;
  unsigned epoch = 0;
  if (threadIdx.x < 32)
    for (int vb = blockIdx.x; vb < nblk; vb += Gn) {
      triinv_diag_body(sm.Xs, vb, k, R, ldr, V, ldv);
      __syncwarp();
    }
  for (int s = TB; s < k; s *= 2) {
    const int nt = s / TB, pairs = (k + 2 * s - 1) / (2 * s);
    for (int phase = 0; phase < 2; ++phase) {
      if (!tri_grid_barrier(bar, ++epoch * Gn, &s_fail)) return;
      for (int t = blockIdx.x; t < pairs * nt * nt; t += Gn) {
        __syncthreads();
        triinv_join_body(sm, t, k, s, phase, R, ldr, V, ldv, tmp, ldt);
      }
    }
  }
}

}  // namespace

// Rinv (k x k, ld ldx) <- R^{-1}.  Rinv's strictly lower triangle is left as the caller set it (zeros).
int bra_tri_inverse_upper(bra_ctx* ctx, int k, const double* R, int64_t ldr, double* Rinv, int64_t ldx) {
  if (k <= 0) return BRA_OK;
  const int nblk = (k + TB - 1) / TB;
  BRA_CUDA(ctx->ws_tritmp().reserve((size_t)k * k * 8 + 256));
  double* tmp = ctx->ws_tritmp().as<double>();
  unsigned* bar = reinterpret_cast<unsigned*>(tmp + (size_t)k * k);
  BRA_CUDA(cudaMemsetAsync(bar, 0, 4, ctx->stream));
  // widest level: ceil(k / 2s) pairs of (s/32)^2 tiles; half the SMs at most (the other lane may run its own cooperative grid)
  int grid = nblk;
  for (int s = TB; s < k; s *= 2) {
    const int nt = s / TB, pairs = (k + 2 * s - 1) / (2 * s);
    if (pairs * nt * nt > grid) grid = pairs * nt * nt;
  }
  const int cap = ctx->num_sms / 2 > 0 ? ctx->num_sms / 2 : 1;
  if (grid > cap) grid = cap;
  int64_t ldt = k;
  void* args[] = {(void*)&k, (void*)&R, (void*)&ldr, (void*)&Rinv, (void*)&ldx, (void*)&tmp, (void*)&ldt, (void*)&bar};
  BRA_CUDA(cudaLaunchCooperativeKernel((void*)triinv_fused_kernel, dim3(grid), dim3(256), args, 0, ctx->stream));
  ctx->launches++;
  BRA_CUDA(cudaGetLastError());
  return BRA_OK;
}

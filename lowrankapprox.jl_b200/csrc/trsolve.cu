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

__global__ void __launch_bounds__(256) trsolve_upper_kernel(int k, int64_t nrhs, const double* __restrict__ R,
                                                            int64_t ldr, double* __restrict__ X, int64_t ldx) {
  __shared__ double Rs[TB][TB + 1];   // Rs[r][c] = R[ib*32 + r, jb*32 + c]
  __shared__ double Xs[TB][TB + 1];   // Xs[r][c] = X[jb*32 + r, col0 + c]
  __shared__ double Acc[TB][TB + 1];
  const int tid = threadIdx.x;
  const int tx = tid & 31;            // rhs column within the panel
  const int ty = tid >> 5;            // row group: rows ty*4 .. ty*4+3
  const int64_t col0 = (int64_t)blockIdx.x * TB;
  const int nblk = (k + TB - 1) / TB;

  for (int ib = nblk - 1; ib >= 0; --ib) {
    const int r0 = ib * TB;
    double acc[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int r = r0 + ty * 4 + u;
      acc[u] = (r < k && col0 + tx < nrhs) ? X[r + (col0 + tx) * ldx] : 0.0;
    }
    for (int jb = ib + 1; jb < nblk; ++jb) {
      const int c0 = jb * TB;
      __syncthreads();
      for (int e = tid; e < TB * TB; e += 256) {
        const int rr = e & 31, cc = e >> 5;
        const int r = r0 + rr, c = c0 + cc;
        Rs[rr][cc] = (r < k && c < k) ? R[r + (int64_t)c * ldr] : 0.0;
        // X rows of block jb were finalised earlier by this same CTA
        const int xr = c0 + rr;
        Xs[rr][cc] = (xr < k && col0 + cc < nrhs) ? X[xr + (col0 + cc) * ldx] : 0.0;
      }
      __syncthreads();
#pragma unroll 8
      for (int kk = 0; kk < TB; ++kk) {
        const double x = Xs[kk][tx];
#pragma unroll
        for (int u = 0; u < 4; ++u) acc[u] = fma(-Rs[ty * 4 + u][kk], x, acc[u]);
      }
    }
    __syncthreads();
    // diagonal block: load R_ii, park acc in smem, one warp substitutes (lane = rhs column)
    for (int e = tid; e < TB * TB; e += 256) {
      const int rr = e & 31, cc = e >> 5;
      const int r = r0 + rr, c = r0 + cc;
      Rs[rr][cc] = (r < k && c < k) ? R[r + (int64_t)c * ldr] : (rr == cc ? 1.0 : 0.0);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) Acc[ty * 4 + u][tx] = acc[u];
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

// In place: X (k x nrhs, holds R12 on entry) <- R11^{-1} X
int bra_trsolve_upper(bra_ctx* ctx, int k, int64_t nrhs, const double* R11, int64_t ldr, double* X, int64_t ldx) {
  if (k <= 0 || nrhs <= 0) return BRA_OK;
  const unsigned grid = (unsigned)((nrhs + TB - 1) / TB);
  trsolve_upper_kernel<<<grid, 256, 0, ctx->stream>>>(k, nrhs, R11, ldr, X, ldx);
  ctx->launches++;
  BRA_CUDA(cudaGetLastError());
  return BRA_OK;
}

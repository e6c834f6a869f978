// Spectral norm estimate by randomised power iteration (reference: snorm / snormdiff, src/snorm.jl:14-53):
//   x <- x / ||x||;  x <- (A - L R)' ((A - L R) x)   (Hermitian A, no low-rank part: x <- A x);   s = sqrt(||x||)  (Hermitian: ||x||)
// until |s - s_prev| <= max(atol, s_prev * rtol) or snorm_niter iterations.  snormdiff(A, F) for a factorization F = L R
// (ID: L = A[:, sk], R = [I T] P';  QR: Q, R;  SVD: U diag(S), Vt) never forms F.
// HBM-bound: two passes over A per iteration (8 m n bytes each); the matrix-vector kernels split the long dimension
// over the grid and reduce the partial results in a fixed order (bitwise reproducible).
#include "common.cuh"
#include <cmath>
#include <cstring>

namespace {

constexpr int GV_ROWS = 256;      // rows per CTA of the y = A x kernel
constexpr int GV_SPLIT = 16;      // column splits of y = A x

// part[s][r] = sum_{j in split s} A[r, j] x[j]      (A column-major: threads = consecutive rows, coalesced)
__global__ void __launch_bounds__(GV_ROWS) gemv_n_kernel(const double* __restrict__ A, int64_t lda, int64_t m, int64_t n,
                                                         const double* __restrict__ x, double* __restrict__ part) {
  const int64_t r = (int64_t)blockIdx.x * GV_ROWS + threadIdx.x;
  const int64_t per = (n + gridDim.y - 1) / gridDim.y;
  const int64_t j0 = (int64_t)blockIdx.y * per, j1 = min(n, j0 + per);
  __shared__ double xs[256];
  double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
  for (int64_t jb = j0; jb < j1; jb += 256) {
    __syncthreads();
    if (jb + threadIdx.x < j1) xs[threadIdx.x] = x[jb + threadIdx.x];
    __syncthreads();
    const int cnt = (int)min((int64_t)256, j1 - jb);
    if (r < m) {
      const double* a = A + r + jb * lda;
      int j = 0;
      for (; j + 3 < cnt; j += 4) {
        a0 = fma(a[(int64_t)j * lda], xs[j], a0);
        a1 = fma(a[(int64_t)(j + 1) * lda], xs[j + 1], a1);
        a2 = fma(a[(int64_t)(j + 2) * lda], xs[j + 2], a2);
        a3 = fma(a[(int64_t)(j + 3) * lda], xs[j + 3], a3);
      }
      for (; j < cnt; ++j) a0 = fma(a[(int64_t)j * lda], xs[j], a0);
    }
  }
  if (r < m) part[(int64_t)blockIdx.y * m + r] = (a0 + a1) + (a2 + a3);
}

// y[r] = alpha * sum_s part[s][r] + beta * y[r]
__global__ void gemv_reduce_kernel(const double* __restrict__ part, int splits, int64_t m, double alpha, double beta,
                                   double* __restrict__ y) {
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < m; r += (int64_t)gridDim.x * blockDim.x) {
    double acc = 0.0;
    for (int s = 0; s < splits; ++s) acc += part[(int64_t)s * m + r];
    y[r] = (beta != 0.0) ? fma(alpha, acc, beta * y[r]) : alpha * acc;
  }
}

// y[j] = alpha * A[:, j] . x + beta * y[j]     (one warp per column, coalesced along the column)
__global__ void __launch_bounds__(256) gemv_t_kernel(const double* __restrict__ A, int64_t lda, int64_t m, int64_t n,
                                                     const double* __restrict__ x, double alpha, double beta,
                                                     double* __restrict__ y) {
  const int lane = threadIdx.x & 31;
  for (int64_t j = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5); j < n; j += (int64_t)gridDim.x * 8) {
    const double* a = A + j * lda;
    double a0 = 0.0, a1 = 0.0;
    int64_t r = lane;
    for (; r + 32 < m; r += 64) {
      a0 = fma(a[r], x[r], a0);
      a1 = fma(a[r + 32], x[r + 32], a1);
    }
    if (r < m) a0 = fma(a[r], x[r], a0);
    double acc = a0 + a1;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) y[j] = (beta != 0.0) ? fma(alpha, acc, beta * y[j]) : alpha * acc;
  }
}

// out[0] = ||x||_2 (single CTA, fixed order);  optionally x <- x * scale first
__global__ void __launch_bounds__(1024) nrm2_scale_kernel(double* __restrict__ x, int64_t n, double scale, int do_scale,
                                                          double* __restrict__ out) {
  __shared__ double red[32];
  double ss = 0.0;
  for (int64_t i = threadIdx.x; i < n; i += 1024) {
    double v = x[i];
    if (do_scale) {
      v *= scale;
      x[i] = v;
    }
    ss = fma(v, v, ss);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ss;
  __syncthreads();
  if (threadIdx.x < 32) {
    ss = red[threadIdx.x];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    if (threadIdx.x == 0) out[0] = sqrt(ss);
  }
}

__global__ void snorm_symcheck_kernel(const double* __restrict__ A, int64_t lda, int64_t n, int* __restrict__ flag) {
  for (int64_t j = blockIdx.x; j < n; j += gridDim.x)
    for (int64_t i = j + 1 + threadIdx.x; i < n; i += blockDim.x)
      if (A[i + j * lda] != A[j + i * lda]) *flag = 1;
}

struct Work {
  bra_ctx* ctx;
  double* part;
  int gemv_n(const double* A, int64_t lda, int64_t m, int64_t n, const double* x, double alpha, double beta, double* y) {
    if (m <= 0) return BRA_OK;
    if (n <= 0) {
      if (beta == 0.0) BRA_CUDA(cudaMemsetAsync(y, 0, (size_t)m * 8, ctx->stream));
      return BRA_OK;
    }
    const int splits = (int)(n < 256 * GV_SPLIT ? (n + 255) / 256 : GV_SPLIT);
    dim3 grid((unsigned)((m + GV_ROWS - 1) / GV_ROWS), (unsigned)splits);
    gemv_n_kernel<<<grid, GV_ROWS, 0, ctx->stream>>>(A, lda, m, n, x, part);
    gemv_reduce_kernel<<<(unsigned)((m + 255) / 256 < 1184 ? (m + 255) / 256 : 1184), 256, 0, ctx->stream>>>(part, splits, m, alpha, beta, y);
    ctx->launches += 2;
    BRA_CUDA(cudaGetLastError());
    return BRA_OK;
  }
  int gemv_t(const double* A, int64_t lda, int64_t m, int64_t n, const double* x, double alpha, double beta, double* y) {
    if (n <= 0) return BRA_OK;
    gemv_t_kernel<<<(unsigned)((n + 7) / 8 < 148 * 8 ? (n + 7) / 8 : 148 * 8), 256, 0, ctx->stream>>>(A, lda, m, n, x, alpha, beta, y);
    ctx->launches++;
    BRA_CUDA(cudaGetLastError());
    return BRA_OK;
  }
};

}  // namespace

// snorm(A - L R): A m x n (lda), L m x k (ldl), R k x n (ldr), all DEVICE resident; k = 0: snorm(A).
// x0: optional device start vector (n entries, the reference's crandn(n)); NULL: device Philox keyed by opts->seed.
extern "C" int bra_snorm_f64(bra_ctx* ctx, int64_t m, int64_t n, const double* A, int64_t lda, int64_t k, const double* L,
                             int64_t ldl, const double* R, int64_t ldr, const bra_opts* opts, int64_t niter_max,
                             const double* x0, double* result, int64_t* niter_out) {
  if (!ctx) return -1;
  BRA_CHECK_ARG(m >= 0, 2, "m");
  BRA_CHECK_ARG(n >= 0, 3, "n");
  BRA_CHECK_ARG(A != nullptr || m * n == 0, 4, "A");
  BRA_CHECK_ARG(lda >= (m > 1 ? m : 1), 5, "lda");
  BRA_CHECK_ARG(k >= 0, 6, "k");
  BRA_CHECK_ARG(k == 0 || (L != nullptr && R != nullptr), 7, "L/R");
  BRA_CHECK_ARG(k == 0 || (ldl >= (m > 1 ? m : 1) && ldr >= k), 8, "ldl/ldr");
  BRA_CHECK_ARG(opts != nullptr && result != nullptr, 11, "opts/result");
  *result = 0.0;
  if (niter_out) *niter_out = 0;
  if (m == 0 || n == 0) return BRA_OK;
  BRA_CUDA(cudaSetDevice(ctx->device));
  if (!is_device_ptr(A) || (k > 0 && (!is_device_ptr(L) || !is_device_ptr(R))) || (x0 && !is_device_ptr(x0))) {
    ctx->set_error("bra_snorm_f64 takes device-resident operands");
    return BRA_ERR_UNSUPPORTED;
  }
  // work: xn (n+1), xm (m), z (k), partials (GV_SPLIT * max(m, k)), nrm (1)
  const int64_t mk = m > k ? m : k;
  BRA_CUDA(ctx->scratch.reserve((size_t)(n + 2 + m + k + (int64_t)GV_SPLIT * mk + 8) * 8));
  double* xn = ctx->scratch.as<double>();
  double* xm = xn + ((n + 2) & ~int64_t(1));
  double* z = xm + m;
  double* part = z + k;
  double* dnrm = part + (int64_t)GV_SPLIT * mk;
  Work w{ctx, part};
  bool isherm = false;
  if (k == 0 && m == n) {
    BRA_CUDA(ctx->info.reserve(64));
    BRA_CUDA(cudaMemsetAsync(ctx->info.as<int>() + 14, 0, 4, ctx->stream));
    snorm_symcheck_kernel<<<(unsigned)(n < 148 * 8 ? n : 148 * 8), 256, 0, ctx->stream>>>(A, lda, n, ctx->info.as<int>() + 14);
    ctx->launches++;
    BRA_CUDA(cudaMemcpyAsync(ctx->h_info + 14, ctx->info.as<int>() + 14, 4, cudaMemcpyDeviceToHost, ctx->stream));
    BRA_CUDA(cudaStreamSynchronize(ctx->stream));
    isherm = ctx->h_info[14] == 0;
  }
  int rc;
  if (x0) BRA_CUDA(cudaMemcpyAsync(xn, x0, (size_t)n * 8, cudaMemcpyDeviceToDevice, ctx->stream));
  else if ((rc = bra_fill_randn(ctx, xn, n, opts->seed, 0x736e6f726dULL))) return rc;
  double* hn = reinterpret_cast<double*>(ctx->h_pin);
  auto norm_of = [&](double scale, int do_scale) -> int {
    nrm2_scale_kernel<<<1, 1024, 0, ctx->stream>>>(xn, n, scale, do_scale, dnrm);
    ctx->launches++;
    BRA_CUDA(cudaMemcpyAsync(hn, dnrm, 8, cudaMemcpyDeviceToHost, ctx->stream));
    BRA_CUDA(cudaStreamSynchronize(ctx->stream));
    return BRA_OK;
  };
  if ((rc = norm_of(1.0, 0))) return rc;
  double xnrm = hn[0], s = 1.0, t = 0.0;
  int64_t niter = 0;
  while (s > 0 && std::fabs(s - t) > std::fmax(opts->atol, t * opts->rtol)) {
    if (niter == niter_max) break;                               // iteration limit (src/snorm.jl:26-31)
    ++niter;
    // xn <- xn / xnrm, fused with nothing else to keep the reference's operation order; then the products
    nrm2_scale_kernel<<<1, 1024, 0, ctx->stream>>>(xn, n, 1.0 / xnrm, 1, dnrm);
    ctx->launches++;
    if ((rc = w.gemv_n(A, lda, m, n, xn, 1.0, 0.0, xm))) return rc;                      // xm = A xn
    if (k > 0) {
      if ((rc = w.gemv_n(R, ldr, k, n, xn, 1.0, 0.0, z))) return rc;                     // z = R xn
      if ((rc = w.gemv_n(L, ldl, m, k, z, -1.0, 1.0, xm))) return rc;                    // xm -= L z
    }
    if (isherm) {
      BRA_CUDA(cudaMemcpyAsync(xn, xm, (size_t)n * 8, cudaMemcpyDeviceToDevice, ctx->stream));
    } else {
      if ((rc = w.gemv_t(A, lda, m, n, xm, 1.0, 0.0, xn))) return rc;                    // xn = A' xm
      if (k > 0) {
        if ((rc = w.gemv_t(L, ldl, m, k, xm, 1.0, 0.0, z))) return rc;                   // z = L' xm
        if ((rc = w.gemv_t(R, ldr, k, n, z, -1.0, 1.0, xn))) return rc;                  // xn -= R' z
      }
    }
    if ((rc = norm_of(1.0, 0))) return rc;
    xnrm = hn[0];
    t = s;
    s = isherm ? xnrm : std::sqrt(xnrm);
  }
  *result = s;
  if (niter_out) *niter_out = niter;
  return BRA_OK;
}

// Strong rank-revealing post-processing of the ID: determinant maximisation by column swaps
// (reference: maxdet_swapcols! / maxdet_update!, src/pqr.jl:444-501; findmaxabs, src/util.jl:13-23).
//
// While max |T_ij| > 1 + maxdet_tol: skeleton column i and redundant column j trade places (p[i] <-> p[k+j]) and
// T = R11^{-1} R12 is updated by the Sherman-Morrison formula of the reference,
//     w1 = T[:, j] - e_i,  T[:, j] <- e_i,  w2 = T[i, :]  (so w2[j] = 1),  T <- T - w1 w2' / (1 + w1[i]).
// Only p and T are maintained: on this path Q and R are never taken from the sketch (pqrfact recomputes them from
// A[:, sk], src/pqr.jl:297-305), so the reference's re-triangularisation of R1 (:467-476) has no counterpart here.
//
// HBM-bound: one read of T for the arg-max and one read + write for the rank-1 update per swap (k (n-k) 8 bytes each).
#include "common.cuh"

namespace {

struct MaxIdx {
  double v;
  long long i;
};

// findmaxabs keeps the LAST maximum in column-major order (`t < m && continue`): larger |.| wins, ties go to the
// larger linear index.
__device__ __forceinline__ void take(MaxIdx& a, double v, long long i) {
  if (v > a.v || (v == a.v && i > a.i)) {
    a.v = v;
    a.i = i;
  }
}

__device__ __forceinline__ MaxIdx block_max(MaxIdx a) {
  __shared__ double sv[32];
  __shared__ long long si[32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const double v = __shfl_xor_sync(0xffffffffu, a.v, o);
    const long long i = __shfl_xor_sync(0xffffffffu, a.i, o);
    take(a, v, i);
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) / 32;
  if (lane == 0) {
    sv[warp] = a.v;
    si[warp] = a.i;
  }
  __syncthreads();
  if (warp == 0) {
    a.v = (lane < nw) ? sv[lane] : -1.0;
    a.i = (lane < nw) ? si[lane] : -1;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const double v = __shfl_xor_sync(0xffffffffu, a.v, o);
      const long long i = __shfl_xor_sync(0xffffffffu, a.i, o);
      take(a, v, i);
    }
  }
  return a;      // valid in warp 0
}

__global__ void __launch_bounds__(256) maxabs_partial_kernel(const double* __restrict__ T, int64_t ld, int k,
                                                             int64_t ncols, double* __restrict__ pv,
                                                             long long* __restrict__ pi) {
  MaxIdx a = {-1.0, -1};
  for (int64_t c = blockIdx.x; c < ncols; c += gridDim.x) {
    const double* t = T + c * ld;
    for (int r = threadIdx.x; r < k; r += blockDim.x) take(a, fabs(t[r]), c * (int64_t)k + r);
  }
  a = block_max(a);
  if (threadIdx.x == 0) {
    pv[blockIdx.x] = a.v;
    pi[blockIdx.x] = a.i;
  }
}

__global__ void __launch_bounds__(256) maxabs_final_kernel(const double* __restrict__ pv, const long long* __restrict__ pi,
                                                           int nparts, double* __restrict__ outv,
                                                           long long* __restrict__ outi) {
  MaxIdx a = {-1.0, -1};
  for (int t = threadIdx.x; t < nparts; t += blockDim.x) take(a, pv[t], pi[t]);
  a = block_max(a);
  if (threadIdx.x == 0) {
    outv[0] = a.v;
    outi[0] = a.i;
  }
}

// w1 = T[:, j] - e_i; T[:, j] = e_i; w2 = T[i, :] with w2[j] = 1; p[i] <-> p[k + j]
__global__ void maxdet_prep_kernel(double* __restrict__ T, int64_t ld, int k, int64_t ncols, int i, int64_t j,
                                   double* __restrict__ w1, double* __restrict__ w2, int64_t* __restrict__ jpvt) {
  const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g < ncols) w2[g] = (g == j) ? 1.0 : T[i + g * ld];
  if (g < k) {
    const double x = T[g + j * ld];
    w1[g] = (g == i) ? x - 1.0 : x;
    T[g + j * ld] = (g == i) ? 1.0 : 0.0;
  }
  if (g == 0) {
    const int64_t t = jpvt[i];
    jpvt[i] = jpvt[k + j];
    jpvt[k + j] = t;
  }
}

// T += alpha w1 w2', alpha = -1 / (1 + w1[i])   (BLAS.ger!, src/pqr.jl:493)
__global__ void __launch_bounds__(256) maxdet_ger_kernel(double* __restrict__ T, int64_t ld, int k, int64_t ncols, int i,
                                                         const double* __restrict__ w1, const double* __restrict__ w2) {
  const double alpha = -1.0 / (1.0 + w1[i]);
  for (int64_t c = blockIdx.x; c < ncols; c += gridDim.x) {
    const double t = alpha * w2[c];
    double* col = T + c * ld;
    for (int r = threadIdx.x; r < k; r += blockDim.x) col[r] = fma(t, w1[r], col[r]);
  }
}

}  // namespace

// In place on the device: T (k x ncols, ld), jpvt (k + ncols entries, 1-based).  Returns the number of swaps in *nswaps.
int bra_maxdet_swapcols(bra_ctx* ctx, int k, int64_t ncols, double* T, int64_t ld, int64_t* jpvt, double tol,
                        int64_t niter_max, int64_t* nswaps) {
  *nswaps = 0;
  if (k <= 0 || ncols <= 0) return BRA_OK;
  const int nparts = (int)(ncols < (int64_t)ctx->num_sms * 4 ? ncols : (int64_t)ctx->num_sms * 4);
  BRA_CUDA(ctx->scratch.reserve((size_t)nparts * 16 + 64 + ((size_t)k + (size_t)ncols) * 8));
  double* pv = ctx->scratch.as<double>();
  long long* pi = reinterpret_cast<long long*>(pv + nparts);
  double* outv = reinterpret_cast<double*>(pi + nparts);
  long long* outi = reinterpret_cast<long long*>(outv + 1);
  double* w1 = outv + 8;
  double* w2 = w1 + k;
  double* hv = reinterpret_cast<double*>(ctx->h_pin);
  long long* hi = reinterpret_cast<long long*>(ctx->h_pin + 8);
  int64_t niter = 0;
  for (;;) {
    maxabs_partial_kernel<<<nparts, 256, 0, ctx->stream>>>(T, ld, k, ncols, pv, pi);
    maxabs_final_kernel<<<1, 256, 0, ctx->stream>>>(pv, pi, nparts, outv, outi);
    ctx->launches += 2;
    BRA_CUDA(cudaMemcpyAsync(hv, outv, 16, cudaMemcpyDeviceToHost, ctx->stream));
    BRA_CUDA(cudaStreamSynchronize(ctx->stream));
    if (!(hv[0] > 1.0 + tol)) break;               // Tmax <= 1 + maxdet_tol (also stops on NaN)
    if (niter == niter_max) break;                  // iteration limit (src/pqr.jl:456-461); -1 never matches
    ++niter;
    const int i = (int)(hi[0] % k);
    const int64_t j = hi[0] / k;
    const int64_t span = ncols > k ? ncols : k;
    maxdet_prep_kernel<<<(unsigned)((span + 255) / 256), 256, 0, ctx->stream>>>(T, ld, k, ncols, i, j, w1, w2, jpvt);
    maxdet_ger_kernel<<<(unsigned)(ncols < (int64_t)ctx->num_sms * 8 ? ncols : (int64_t)ctx->num_sms * 8), 256, 0,
                        ctx->stream>>>(T, ld, k, ncols, i, w1, w2);
    ctx->launches += 2;
  }
  BRA_CUDA(cudaGetLastError());
  *nswaps = niter;
  return BRA_OK;
}

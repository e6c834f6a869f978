// pheigfact / pheigvals on the device (reference: src/pheig.jl:276-296, 298-311; pheigrank :322-341).
//   V = idfact(:n, A);  Q R = qr(Matrix(:c, V)) with Matrix(:c, V) = P [I; T'] (n x k);  B = R (A[sk,sk] R');
//   eigen(hermitianize(B)), truncation by pheigrank, vectors = Q * F.vectors.
// Z = [I; T'] is well conditioned, so Q R = Z comes from ONE Cholesky pass on Z'Z = I + T T' and Q is never formed
// (vectors = Z (R^{-1} W), like the right factor of psvdfact).  The k x k symmetric eigenproblem reuses the
// one-sided Jacobi: B J = Y Sigma with B symmetric gives eigenvectors J and eigenvalues as the Rayleigh quotients
// lambda_i = j_i' (B j_i) = <J[:, i], (B J)[:, i]>, taken from the very columns the Jacobi leaves behind.
#include "common.cuh"
#include <algorithm>
#include <cmath>
#include <cstring>
#include <numeric>
#include <string>
#include <vector>

int bra_gather_scale_cols(bra_ctx* ctx, const double* X, int64_t ldx, int64_t rows, int kk, const int* order_dev,
                          const double* scale_dev, double* out, int64_t ldo);
int bra_scatter_cols(bra_ctx* ctx, const double* src, int64_t lds, int64_t rows, int64_t n, const int64_t* jpvt1,
                     double* dst, int64_t ldd);

namespace {

inline int64_t even(int64_t x) { return (x + 1) & ~int64_t(1); }

__global__ void pheig_symcheck_kernel(const double* __restrict__ A, int64_t lda, int64_t n, int* __restrict__ flag) {
  for (int64_t j = blockIdx.x; j < n; j += gridDim.x)
    for (int64_t i = j + 1 + threadIdx.x; i < n; i += blockDim.x)
      if (A[i + j * lda] != A[j + i * lda]) *flag = 1;
}

// Ask[i, j] = A[sk_i, sk_j]
__global__ void gather_sub_kernel(const double* __restrict__ A, int64_t lda, const int64_t* __restrict__ sk1, int k,
                                  double* __restrict__ out, int64_t ldo) {
  for (int j = blockIdx.x; j < k; j += gridDim.x) {
    const double* a = A + (sk1[j] - 1) * lda;
    for (int i = threadIdx.x; i < k; i += blockDim.x) out[i + (int64_t)j * ldo] = a[sk1[i] - 1];
  }
}

// B <- (B + B') / 2  (hermitianize!, src/util.jl)
__global__ void hermitianize_kernel(double* __restrict__ Bm, int64_t ld, int k) {
  for (int j = blockIdx.x; j < k; j += gridDim.x)
    for (int i = j + 1 + threadIdx.x; i < k; i += blockDim.x) {
      const double v = 0.5 * (Bm[i + (int64_t)j * ld] + Bm[j + (int64_t)i * ld]);
      Bm[i + (int64_t)j * ld] = v;
      Bm[j + (int64_t)i * ld] = v;
    }
}

// lam[i] = <J[:, i], X[:, i]>  (Rayleigh quotient j_i' B j_i, X = B J)
__global__ void rayleigh_kernel(const double* __restrict__ X, const double* __restrict__ J, int64_t ld, int k,
                                double* __restrict__ lam) {
  const int i = blockIdx.x;
  double a = 0.0;
  for (int r = threadIdx.x; r < k; r += 128) a = fma(X[r + (int64_t)i * ld], J[r + (int64_t)i * ld], a);
  __shared__ double red[4];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = a;
  __syncthreads();
  if (threadIdx.x == 0) lam[i] = (red[0] + red[1]) + (red[2] + red[3]);
}

}  // namespace

extern "C" int bra_pheigfact_f64(bra_ctx* ctx, int64_t n, const double* A, int64_t lda, const bra_opts* opts,
                                 const bra_rand* rnd) {
  if (!ctx) return -1;
  int rc = bra_check_fact_args(ctx, 'n', n, n, A, lda, opts);          // checksquare: one dimension argument
  if (rc) return rc;
  BRA_CUDA(cudaSetDevice(ctx->device));
  const double* dA = A;
  int64_t dlda = lda;
  if (!is_device_ptr(A)) {
    dlda = even(n);
    BRA_CUDA(ctx->A_stage.reserve((size_t)dlda * (n > 0 ? n : 1) * 8));
    if (n > 0)
      BRA_CUDA(cudaMemcpy2DAsync(ctx->A_stage.p, (size_t)dlda * 8, A, (size_t)lda * 8, (size_t)n * 8, (size_t)n,
                                 cudaMemcpyDefault, ctx->stream));
    dA = ctx->A_stage.as<double>();
  }
  // !ishermitian(A) && error("matrix must be Hermitian")   (src/pheig.jl:279)
  if (n > 1) {
    BRA_CUDA(ctx->info.reserve(64));
    BRA_CUDA(cudaMemsetAsync(ctx->info.as<int>() + 14, 0, 4, ctx->stream));
    pheig_symcheck_kernel<<<(unsigned)(n < 148 * 8 ? n : 148 * 8), 256, 0, ctx->stream>>>(dA, dlda, n, ctx->info.as<int>() + 14);
    ctx->launches++;
    BRA_CUDA(cudaMemcpyAsync(ctx->h_info + 14, ctx->info.as<int>() + 14, 4, cudaMemcpyDeviceToHost, ctx->stream));
    BRA_CUDA(cudaStreamSynchronize(ctx->stream));
    if (ctx->h_info[14] != 0) {
      ctx->set_error("invalid argument 3: matrix must be Hermitian");
      return -3;
    }
  }
  rc = bra_sketchfact_core(ctx, 'n', n, n, dA, dlda, opts, rnd);       // V = idfact(:n, A, opts)
  if (rc) return rc;
  FactResult& res = ctx->res;
  const int64_t k = res.k;
  res.ksvd = 0;
  res.svd_m = n;
  res.svd_n = n;
  res.have_svd = true;
  if (k == 0) return BRA_OK;
  const int gemm_tag = ctx->gemm_tag;
  ctx->gemm_tag = BRA_PROF_TAILGEMM;
  struct TagRestore {
    bra_ctx* c;
    int t;
    ~TagRestore() { c->gemm_tag = t; }
  } restore{ctx, gemm_tag};

  // Z = [I; T'] (n x k, pivoted row order), R_z from one Cholesky pass on Z'Z
  const int64_t ldz = even(n), ldj = even(k);
  BRA_CUDA(ctx->Z.reserve((size_t)ldz * k * 8));
  double* Z = ctx->Z.as<double>();
  if ((rc = bra_set_identity(ctx, (int)k, Z, ldz))) return rc;
  if (n > k && (rc = bra_transpose(ctx, ctx->T.as<double>(), res.ldT, k, n - k, Z + k, ldz))) return rc;
  BRA_CUDA(ctx->W.reserve((size_t)7 * ldj * k * 8 + 64));
  double* Rz = ctx->W.as<double>();
  double* Bm = Rz + (size_t)ldj * k;
  double* J = Bm + (size_t)ldj * k;
  double* Ask = J + (size_t)ldj * k;
  double* Tmp = Ask + (size_t)ldj * k;
  double* Rinv = Tmp + (size_t)ldj * k;
  double* Wsel = Rinv + (size_t)ldj * k;
  BRA_CUDA(ctx->G.reserve((size_t)k * k * 8));
  if ((rc = bra_chol_status_reset(ctx))) return rc;
  if ((rc = bra_gemm_tn(ctx, Z, ldz, k, n, Z, ldz, k, ctx->G.as<double>(), k))) return rc;
  if ((rc = bra_cholesky_upper(ctx, (int)k, ctx->G.as<double>(), k, Rz, ldj))) return rc;
  // B = R_z (A[sk, sk] R_z')
  gather_sub_kernel<<<(unsigned)(k < 148 * 8 ? k : 148 * 8), 128, 0, ctx->stream>>>(dA, dlda, ctx->jpvt.as<int64_t>(), (int)k, Ask, ldj);
  ctx->launches++;
  // Tmp[i, j] = sum_t Ask[i, t] Rz[j, t]
  if ((rc = bra_gemm_generic(ctx, Ask, 1, ldj, Rz, ldj, 1, k, k, k, Tmp, ldj))) return rc;
  // Bm[i, j] = sum_t Rz[i, t] Tmp[t, j]
  if ((rc = bra_gemm_generic(ctx, Rz, 1, ldj, Tmp, 1, ldj, k, k, k, Bm, ldj))) return rc;
  hermitianize_kernel<<<(unsigned)(k < 148 * 4 ? k : 148 * 4), 128, 0, ctx->stream>>>(Bm, ldj, (int)k);
  ctx->launches++;
  BRA_CUDA(cudaGetLastError());
  if ((rc = bra_chol_status(ctx))) return rc;

  // eigen!(B): one-sided Jacobi, B J = Y Sigma; eigenvalues = Rayleigh quotients
  std::vector<double> sig((size_t)k), lam((size_t)k);
  std::vector<int> order((size_t)k);
  BRA_CUDA(ctx->S.reserve((size_t)2 * k * 8));
  rc = bra_jacobi_svd(ctx, (int)k, Bm, ldj, J, ldj, sig.data(), order.data());
  if (rc) return rc;
  double* lam_dev = ctx->S.as<double>() + k;
  rayleigh_kernel<<<(unsigned)k, 128, 0, ctx->stream>>>(Bm, J, ldj, (int)k, lam_dev);
  ctx->launches++;
  if ((size_t)k * 8 > BRA_HPIN_BYTES) {
    ctx->set_error("pheigfact: k too large for the pinned read-back buffer");
    return BRA_ERR_UNSUPPORTED;
  }
  BRA_CUDA(cudaMemcpyAsync(ctx->h_pin, lam_dev, (size_t)k * 8, cudaMemcpyDeviceToHost, ctx->stream));
  BRA_CUDA(cudaStreamSynchronize(ctx->stream));
  std::memcpy(lam.data(), ctx->h_pin, (size_t)k * 8);
  const double smax = *std::max_element(sig.begin(), sig.end());
  for (int64_t i = 0; i < k; ++i)
    if (std::fabs(std::fabs(lam[(size_t)i]) - sig[(size_t)i]) > 1e-9 * smax) {
      // a +lambda / -lambda pair of equal magnitude: the singular subspace does not separate the two eigenvectors
      ctx->set_error("pheigfact: eigenvalues of equal magnitude and opposite sign are not separated by the Jacobi core");
      return BRA_ERR_UNSUPPORTED;
    }
  // ascending order like eigen!, then pheigrank (src/pheig.jl:322-341)
  std::vector<int> asc((size_t)k);
  std::iota(asc.begin(), asc.end(), 0);
  std::stable_sort(asc.begin(), asc.end(), [&](int a, int b) { return lam[(size_t)a] < lam[(size_t)b]; });
  std::vector<double> w((size_t)k);
  for (int64_t i = 0; i < k; ++i) w[(size_t)i] = lam[(size_t)asc[(size_t)i]];
  const double wmax = std::max(std::fabs(w[0]), std::fabs(w[(size_t)k - 1]));
  const double ptol = std::max(opts->atol, opts->rtol * wmax);
  const int64_t nneg = std::lower_bound(w.begin(), w.end(), 0.0) - w.begin();          // first(idx) - 1
  const int64_t npos = w.end() - std::upper_bound(w.begin(), w.end(), 0.0);             // n - last(idx)
  auto rank1 = [&](int64_t cnt, bool from_top) {
    int64_t kk = opts->rank >= 0 ? std::min<int64_t>(opts->rank, cnt) : cnt;
    for (int64_t i = 1; i < kk; ++i) {
      const double v = from_top ? w[(size_t)(k - 1 - i)] : w[(size_t)i];
      if (std::fabs(v) <= ptol) return i;
    }
    return kk;
  };
  const int64_t kn = rank1(nneg, false), kp = rank1(npos, true);
  std::vector<int> sel;
  std::vector<double> vals;
  if (kn + kp < k) {
    for (int64_t i = 0; i < kn; ++i) sel.push_back(asc[(size_t)i]);
    for (int64_t i = k - kp; i < k; ++i) sel.push_back(asc[(size_t)i]);
  } else {
    sel = asc;
  }
  for (int s : sel) vals.push_back(lam[(size_t)s]);
  const int64_t kk = (int64_t)sel.size();
  res.ksvd = kk;
  if (kk == 0) return BRA_OK;

  // vectors = Q_z J[:, sel] = Z (R_z^{-1} J[:, sel]);  pivoted rows -> original rows
  BRA_CUDA(ctx->aux_in1.reserve((size_t)k * 4));
  {
    int* ho = reinterpret_cast<int*>(ctx->h_pin);
    double* hv = reinterpret_cast<double*>(ctx->h_pin + (((size_t)k * 4 + 63) & ~size_t(63)));
    std::memcpy(ho, sel.data(), (size_t)kk * 4);
    std::memcpy(hv, vals.data(), (size_t)kk * 8);
    BRA_CUDA(cudaMemcpyAsync(ctx->aux_in1.p, ho, (size_t)kk * 4, cudaMemcpyHostToDevice, ctx->stream));
    BRA_CUDA(cudaMemcpyAsync(ctx->S.p, hv, (size_t)kk * 8, cudaMemcpyHostToDevice, ctx->stream));
  }
  if ((rc = bra_gather_scale_cols(ctx, J, ldj, k, (int)kk, ctx->aux_in1.as<int>(), nullptr, Wsel, ldj))) return rc;
  if ((rc = bra_set_identity(ctx, (int)k, Rinv, ldj))) return rc;
  if ((rc = bra_tri_inverse_upper(ctx, (int)k, Rz, ldj, Rinv, ldj))) return rc;
  if ((rc = bra_gemm_generic(ctx, Rinv, 1, ldj, Wsel, 1, ldj, k, kk, k, Tmp, ldj))) return rc;        // Yh = R_z^{-1} Wsel
  const int64_t ldv = even(kk);
  BRA_CUDA(ctx->B2.reserve((size_t)ldv * n * 8));
  BRA_CUDA(ctx->scratch3.reserve((size_t)ldv * n * 8 + 64));
  BRA_CUDA(ctx->U.reserve((size_t)n * kk * 8 + 64));
  if ((rc = bra_transpose(ctx, Tmp, ldj, k, kk, ctx->B2.as<double>(), ldv))) return rc;                 // first k columns: Yh'
  if (n > k && (rc = bra_gemm_tn(ctx, Tmp, ldj, kk, k, ctx->T.as<double>(), res.ldT, n - k,
                                 ctx->B2.as<double>() + (size_t)ldv * k, ldv)))
    return rc;
  if ((rc = bra_scatter_cols(ctx, ctx->B2.as<double>(), ldv, kk, n, ctx->jpvt.as<int64_t>(), ctx->scratch3.as<double>(), ldv)))
    return rc;
  if ((rc = bra_transpose(ctx, ctx->scratch3.as<double>(), ldv, kk, n, ctx->U.as<double>(), n))) return rc;   // n x kk
  BRA_CUDA(cudaStreamSynchronize(ctx->stream));
  return BRA_OK;
}

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

// pheigorth! (src/pheig.jl:342-364): runs of adjacent eigenvalues with symrelerr(va, vb) <= orthtol (src/util.jl:118) are
// re-orthogonalised by the reference's own sweep -- for i in the run, for j > i in the run: v_j -= <v_i, v_j> v_i, no
// normalisation.  One CTA; values ascending, vectors = columns of V (rows x kk).
__global__ void __launch_bounds__(256) pheigorth_kernel(const double* __restrict__ vals, double* __restrict__ V, int64_t ld,
                                                        int rows, int kk, double orthtol) {
  __shared__ double red[8];
  __shared__ double s_dot;
  const int tid = threadIdx.x;
  int a = 0;
  while (a < kk) {
    const double va = vals[a];
    int b = a + 1;
    while (b < kk) {
      const double vb = vals[b];
      if (2.0 * fabs((va - vb) / (va + vb)) > orthtol) break;      // NaN (0/0) compares false, like the reference
      ++b;
    }
    --b;
    for (int i = a; i <= b; ++i) {
      const double* vi = V + (int64_t)i * ld;
      for (int j = i + 1; j <= b; ++j) {
        double* vj = V + (int64_t)j * ld;
        double d = 0.0;
        for (int r = tid; r < rows; r += 256) d = fma(vi[r], vj[r], d);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
        if ((tid & 31) == 0) red[tid >> 5] = d;
        __syncthreads();
        if (tid == 0) s_dot = ((red[0] + red[1]) + (red[2] + red[3])) + ((red[4] + red[5]) + (red[6] + red[7]));
        __syncthreads();
        const double dd = s_dot;
        for (int r = tid; r < rows; r += 256) vj[r] = fma(-dd, vi[r], vj[r]);
        __syncthreads();
      }
    }
    a = b + 1;
  }
}

// cyclic two-sided Jacobi for a tiny symmetric matrix H (c x c, row-major on the host): H = Z diag(w) Z'
void host_sym_eig(int c, std::vector<double>& H, std::vector<double>& Z, std::vector<double>& w) {
  Z.assign((size_t)c * c, 0.0);
  for (int i = 0; i < c; ++i) Z[(size_t)i * c + i] = 1.0;
  for (int sweep = 0; sweep < 60; ++sweep) {
    double off = 0.0, dia = 0.0;
    for (int i = 0; i < c; ++i)
      for (int j = 0; j < c; ++j) (i == j ? dia : off) += H[(size_t)i * c + j] * H[(size_t)i * c + j];
    if (off <= 1e-32 * dia || off == 0.0) break;
    for (int p = 0; p < c; ++p)
      for (int q = p + 1; q < c; ++q) {
        const double apq = H[(size_t)p * c + q];
        if (apq == 0.0) continue;
        const double theta = (H[(size_t)q * c + q] - H[(size_t)p * c + p]) / (2.0 * apq);
        const double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
        const double cs = 1.0 / std::sqrt(t * t + 1.0), sn = t * cs;
        for (int r = 0; r < c; ++r) {          // H <- H G
          const double hp = H[(size_t)r * c + p], hq = H[(size_t)r * c + q];
          H[(size_t)r * c + p] = cs * hp - sn * hq;
          H[(size_t)r * c + q] = sn * hp + cs * hq;
        }
        for (int r = 0; r < c; ++r) {          // H <- G' H
          const double hp = H[(size_t)p * c + r], hq = H[(size_t)q * c + r];
          H[(size_t)p * c + r] = cs * hp - sn * hq;
          H[(size_t)q * c + r] = sn * hp + cs * hq;
        }
        for (int r = 0; r < c; ++r) {          // Z <- Z G
          const double zp = Z[(size_t)r * c + p], zq = Z[(size_t)r * c + q];
          Z[(size_t)r * c + p] = cs * zp - sn * zq;
          Z[(size_t)r * c + q] = sn * zp + cs * zq;
        }
      }
  }
  w.resize((size_t)c);
  for (int i = 0; i < c; ++i) w[(size_t)i] = H[(size_t)i * c + i];
}

}  // namespace

// eigen!(Hermitian(B)) for a real symmetric k x k B on the device (src/pheig.jl:283, src/cur.jl:103):
// one-sided Jacobi B J = Y Sigma -- for a symmetric B the columns of J are eigenvectors and the Rayleigh quotients
// lambda_i = <J[:, i], (B J)[:, i]> the signed eigenvalues -- EXCEPT inside a group of (nearly) equal |lambda|, where
// the singular vectors are an arbitrary basis of the group's invariant subspace (a +lambda / -lambda pair above all).
// Such groups (adjacent sigma within 1e-6 relative) are resolved by a Rayleigh-Ritz step on the group's c columns:
// H = J_c' (B J_c) (c x c, on the host), H = Z diag(w) Z', J_c <- J_c Z.  On return Bm holds B J (for the rotated
// groups: of the new J_c), J the eigenvectors, lam the eigenvalues (unsorted).
int bra_sym_eigen(bra_ctx* ctx, int k, double* Bm, int64_t ld, double* J, std::vector<double>& lam) {
  std::vector<double> sig((size_t)k);
  std::vector<int> order((size_t)k);
  lam.assign((size_t)k, 0.0);
  BRA_CUDA(ctx->S.reserve((size_t)2 * k * 8));
  int rc = bra_jacobi_svd(ctx, k, Bm, ld, J, ld, sig.data(), order.data());
  if (rc) return rc;
  double* lam_dev = ctx->S.as<double>() + k;
  rayleigh_kernel<<<(unsigned)k, 128, 0, ctx->stream>>>(Bm, J, ld, k, lam_dev);
  ctx->launches++;
  if ((size_t)k * 8 > BRA_HPIN_BYTES) {
    ctx->set_error("symmetric eigensolver: k too large for the pinned read-back buffer");
    return BRA_ERR_UNSUPPORTED;
  }
  BRA_CUDA(cudaMemcpyAsync(ctx->h_pin, lam_dev, (size_t)k * 8, cudaMemcpyDeviceToHost, ctx->stream));
  BRA_CUDA(cudaStreamSynchronize(ctx->stream));
  std::memcpy(lam.data(), ctx->h_pin, (size_t)k * 8);
  // groups of nearly equal singular values, in descending order of sigma
  const double smax = sig[(size_t)order[0]];
  std::vector<double> Jc, Xc, H, Z, w, Jn, Xn;
  for (int a = 0; a < k;) {
    int b = a;
    while (b + 1 < k && sig[(size_t)order[(size_t)b]] - sig[(size_t)order[(size_t)b + 1]] <= 1e-6 * sig[(size_t)order[(size_t)b]] &&
           sig[(size_t)order[(size_t)b]] > 1e-14 * smax)
      ++b;
    const int c = b - a + 1;
    if (c >= 2) {
      Jc.resize((size_t)k * c);
      Xc.resize((size_t)k * c);
      for (int t = 0; t < c; ++t) {
        const int col = order[(size_t)(a + t)];
        BRA_CUDA(cudaMemcpyAsync(Jc.data() + (size_t)t * k, J + (int64_t)col * ld, (size_t)k * 8, cudaMemcpyDeviceToHost, ctx->stream));
        BRA_CUDA(cudaMemcpyAsync(Xc.data() + (size_t)t * k, Bm + (int64_t)col * ld, (size_t)k * 8, cudaMemcpyDeviceToHost, ctx->stream));
      }
      BRA_CUDA(cudaStreamSynchronize(ctx->stream));
      H.assign((size_t)c * c, 0.0);
      for (int i = 0; i < c; ++i)
        for (int j = 0; j < c; ++j) {
          double d = 0.0;
          for (int r = 0; r < k; ++r) d += Jc[(size_t)i * k + r] * Xc[(size_t)j * k + r];
          H[(size_t)i * c + j] = d;
        }
      for (int i = 0; i < c; ++i)
        for (int j = i + 1; j < c; ++j) H[(size_t)i * c + j] = H[(size_t)j * c + i] = 0.5 * (H[(size_t)i * c + j] + H[(size_t)j * c + i]);
      host_sym_eig(c, H, Z, w);
      Jn.assign((size_t)k * c, 0.0);
      Xn.assign((size_t)k * c, 0.0);
      for (int t = 0; t < c; ++t)
        for (int u = 0; u < c; ++u) {
          const double z = Z[(size_t)u * c + t];
          for (int r = 0; r < k; ++r) {
            Jn[(size_t)t * k + r] += z * Jc[(size_t)u * k + r];
            Xn[(size_t)t * k + r] += z * Xc[(size_t)u * k + r];
          }
        }
      for (int t = 0; t < c; ++t) {
        const int col = order[(size_t)(a + t)];
        lam[(size_t)col] = w[(size_t)t];
        BRA_CUDA(cudaMemcpyAsync(J + (int64_t)col * ld, Jn.data() + (size_t)t * k, (size_t)k * 8, cudaMemcpyHostToDevice, ctx->stream));
        BRA_CUDA(cudaMemcpyAsync(Bm + (int64_t)col * ld, Xn.data() + (size_t)t * k, (size_t)k * 8, cudaMemcpyHostToDevice, ctx->stream));
      }
      BRA_CUDA(cudaStreamSynchronize(ctx->stream));      // the host vectors are reused by the next group
    }
    a = b + 1;
  }
  return BRA_OK;
}

int bra_pheigorth(bra_ctx* ctx, const double* vals_dev, double* V, int64_t ld, int rows, int kk, double orthtol) {
  if (kk <= 1 || rows <= 0) return BRA_OK;
  pheigorth_kernel<<<1, 256, 0, ctx->stream>>>(vals_dev, V, ld, rows, kk, orthtol);
  ctx->launches++;
  BRA_CUDA(cudaGetLastError());
  return BRA_OK;
}

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

  // eigen!(hermitianize!(B)) (src/pheig.jl:283)
  std::vector<double> lam;
  if ((rc = bra_sym_eigen(ctx, (int)k, Bm, ldj, J, lam))) return rc;
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
  if ((rc = bra_pheigorth(ctx, ctx->S.as<double>(), Wsel, ldj, (int)k, (int)kk, opts->pheig_orthtol))) return rc;   // :293
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


namespace {
// out[i, j] = A[rows_i, cols_j] (1-based index vectors on the device); rows1 == nullptr: all rows, cols1 == nullptr: all
// columns -- one kernel for C = A[:, cols], R = A[rows, :] and the k x k core
__global__ void gather_rc_kernel(const double* __restrict__ A, int64_t lda, const int64_t* __restrict__ rows1, int64_t nr,
                                 const int64_t* __restrict__ cols1, int64_t nc, double* __restrict__ out, int64_t ldo) {
  for (int64_t j = blockIdx.x; j < nc; j += gridDim.x) {
    const double* a = A + (cols1 ? cols1[j] - 1 : j) * lda;
    for (int64_t i = threadIdx.x; i < nr; i += blockDim.x) out[i + j * ldo] = a[rows1 ? rows1[i] - 1 : i];
  }
}
__global__ void reciprocal_kernel(double* __restrict__ x, int k) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < k) x[i] = 1.0 / x[i];
}
}  // namespace

// CUR(A, U::CURPackedU) / HermCUR(A, U) (src/cur.jl:85-109): see include/brapprox.h
extern "C" int bra_cur_f64(bra_ctx* ctx, int64_t m, int64_t n, const double* A, int64_t lda, int64_t k,
                           const int64_t* rows1, const int64_t* cols1, int hermitian) {
  if (!ctx) return -1;
  BRA_CHECK_ARG(m >= 0, 2, "m");
  BRA_CHECK_ARG(n >= 0, 3, "n");
  BRA_CHECK_ARG(A != nullptr || m * n == 0, 4, "A");
  BRA_CHECK_ARG(lda >= (m > 1 ? m : 1), 5, "lda");
  BRA_CHECK_ARG(k >= 0 && k <= (m < n ? m : n), 6, "k");
  BRA_CHECK_ARG(hermitian || rows1 != nullptr || k == 0, 7, "rows");
  BRA_CHECK_ARG(cols1 != nullptr || k == 0, 8, "cols");
  BRA_CHECK_ARG(!hermitian || m == n, 9, "a Hermitian CUR needs a square matrix");
  BRA_CUDA(cudaSetDevice(ctx->device));
  FactResult& res = ctx->res;
  res = FactResult();
  res.m = m;
  res.n = n;
  res.k = k;
  res.ksvd = k;
  res.svd_m = k;
  res.svd_n = k;
  res.have_Q = true;
  res.have_R = !hermitian;
  res.have_svd = true;
  if (k == 0) return BRA_OK;
  // index vectors: bounds-checked on the host (BoundsError in the reference), then staged
  std::vector<int64_t> hidx((size_t)2 * k);
  const int64_t* hr = hermitian ? cols1 : rows1;
  if (is_device_ptr(cols1)) {
    BRA_CUDA(cudaMemcpy(hidx.data() + k, cols1, (size_t)k * 8, cudaMemcpyDeviceToHost));
    BRA_CUDA(cudaMemcpy(hidx.data(), hr, (size_t)k * 8, cudaMemcpyDeviceToHost));
  } else {
    std::memcpy(hidx.data() + k, cols1, (size_t)k * 8);
    std::memcpy(hidx.data(), hr, (size_t)k * 8);
  }
  for (int64_t i = 0; i < k; ++i) {
    BRA_CHECK_ARG(hidx[(size_t)i] >= 1 && hidx[(size_t)i] <= m, 7, "row index out of bounds");
    BRA_CHECK_ARG(hidx[(size_t)(k + i)] >= 1 && hidx[(size_t)(k + i)] <= n, 8, "column index out of bounds");
  }
  BRA_CUDA(ctx->aux_in2.reserve((size_t)2 * k * 8));
  BRA_CUDA(cudaMemcpyAsync(ctx->aux_in2.p, hidx.data(), (size_t)2 * k * 8, cudaMemcpyHostToDevice, ctx->stream));
  BRA_CUDA(cudaStreamSynchronize(ctx->stream));           // hidx is a stack-owned vector
  const int64_t* drows = ctx->aux_in2.as<int64_t>();
  const int64_t* dcols = drows + k;
  const double* dA = A;
  int64_t dlda = lda;
  if (!is_device_ptr(A)) {
    dlda = even(m);
    BRA_CUDA(ctx->A_stage.reserve((size_t)dlda * (n > 0 ? n : 1) * 8));
    BRA_CUDA(cudaMemcpy2DAsync(ctx->A_stage.p, (size_t)dlda * 8, A, (size_t)lda * 8, (size_t)m * 8, (size_t)n,
                               cudaMemcpyDefault, ctx->stream));
    dA = ctx->A_stage.as<double>();
  }
  const int64_t ldq = even(m), ldj = even(k);
  BRA_CUDA(ctx->Q.reserve((size_t)ldq * k * 8));
  const unsigned gk = (unsigned)(k < 148 * 8 ? k : 148 * 8);
  gather_rc_kernel<<<gk, 256, 0, ctx->stream>>>(dA, dlda, nullptr, m, dcols, k, ctx->Q.as<double>(), ldq);     // C = A[:, cols]
  ctx->launches++;
  if (!hermitian) {
    BRA_CUDA(ctx->Rfull.reserve((size_t)k * n * 8));
    gather_rc_kernel<<<(unsigned)(n < 148 * 8 ? n : 148 * 8), 128, 0, ctx->stream>>>(dA, dlda, drows, k, nullptr, n,
                                                                                  ctx->Rfull.as<double>(), k);   // R = A[rows, :]
    ctx->launches++;
  }
  BRA_CUDA(ctx->W.reserve((size_t)3 * ldj * k * 8 + 64));
  double* X = ctx->W.as<double>();
  double* J = X + (size_t)ldj * k;
  double* Tmp = J + (size_t)ldj * k;
  gather_rc_kernel<<<gk, 128, 0, ctx->stream>>>(dA, dlda, drows, k, dcols, k, X, ldj);                          // C[rows, :]
  ctx->launches++;
  BRA_CUDA(cudaGetLastError());
  BRA_CUDA(ctx->S.reserve((size_t)2 * k * 8));
  BRA_CUDA(ctx->U.reserve((size_t)k * k * 8 + 64));
  BRA_CUDA(ctx->Vt.reserve((size_t)k * k * 8 + 64));
  BRA_CUDA(ctx->aux_in1.reserve((size_t)k * 4));
  if ((size_t)k * 12 + 128 > BRA_HPIN_BYTES) {
    ctx->set_error("CUR: k too large for the pinned staging buffer");
    return BRA_ERR_UNSUPPORTED;
  }
  int rc;
  int* ho = reinterpret_cast<int*>(ctx->h_pin);
  double* hv = reinterpret_cast<double*>(ctx->h_pin + (((size_t)k * 4 + 63) & ~size_t(63)));
  if (hermitian) {
    hermitianize_kernel<<<(unsigned)(k < 148 * 4 ? k : 148 * 4), 128, 0, ctx->stream>>>(X, ldj, (int)k);          // Hermitian(.)
    ctx->launches++;
    std::vector<double> lam;
    if ((rc = bra_sym_eigen(ctx, (int)k, X, ldj, J, lam))) return rc;
    std::vector<int> asc((size_t)k);
    std::iota(asc.begin(), asc.end(), 0);
    std::stable_sort(asc.begin(), asc.end(), [&](int a, int b) { return lam[(size_t)a] < lam[(size_t)b]; });
    for (int64_t i = 0; i < k; ++i) {
      ho[i] = asc[(size_t)i];
      hv[i] = 1.0 / lam[(size_t)asc[(size_t)i]];                                                                // 1 ./ F.values
    }
    BRA_CUDA(cudaMemcpyAsync(ctx->aux_in1.p, ho, (size_t)k * 4, cudaMemcpyHostToDevice, ctx->stream));
    BRA_CUDA(cudaMemcpyAsync(ctx->S.p, hv, (size_t)k * 8, cudaMemcpyHostToDevice, ctx->stream));
    if ((rc = bra_gather_scale_cols(ctx, J, ldj, k, (int)k, ctx->aux_in1.as<int>(), nullptr, ctx->U.as<double>(), k))) return rc;
  } else {
    // X J = U Sigma: V = J[:, order], U = X[:, order] / sigma
    std::vector<double> sig((size_t)k);
    std::vector<int> order((size_t)k);
    if ((rc = bra_jacobi_svd(ctx, (int)k, X, ldj, J, ldj, sig.data(), order.data()))) return rc;
    for (int64_t i = 0; i < k; ++i) {
      ho[i] = order[(size_t)i];
      hv[i] = 1.0 / sig[(size_t)order[(size_t)i]];                                                              // 1 ./ sigma
    }
    BRA_CUDA(cudaMemcpyAsync(ctx->aux_in1.p, ho, (size_t)k * 4, cudaMemcpyHostToDevice, ctx->stream));
    // bra_jacobi_svd left the unsorted column norms in ctx->S[0:k]: scale by them, then overwrite with the sorted 1/sigma
    if ((rc = bra_gather_scale_cols(ctx, X, ldj, k, (int)k, ctx->aux_in1.as<int>(), ctx->S.as<double>(), Tmp, ldj))) return rc;
    if ((rc = bra_gather_scale_cols(ctx, J, ldj, k, (int)k, ctx->aux_in1.as<int>(), nullptr, ctx->U.as<double>(), k))) return rc;
    if ((rc = bra_transpose(ctx, Tmp, ldj, k, k, ctx->Vt.as<double>(), k))) return rc;                           // U'
    BRA_CUDA(cudaMemcpyAsync(ctx->S.p, hv, (size_t)k * 8, cudaMemcpyHostToDevice, ctx->stream));
  }
  BRA_CUDA(cudaStreamSynchronize(ctx->stream));
  return BRA_OK;
}


// test hook: pheigorth! on host arrays (vals[kk] ascending, V rows x kk with leading dimension ld), in place
extern "C" int bra_debug_pheigorth(bra_ctx* ctx, const double* vals, double* V, int64_t ld, int rows, int kk, double orthtol) {
  if (!ctx) return -1;
  BRA_CUDA(cudaSetDevice(ctx->device));
  BRA_CUDA(ctx->scratch.reserve((size_t)(kk > 0 ? kk : 1) * 8));
  BRA_CUDA(ctx->scratch2.reserve((size_t)ld * (kk > 0 ? kk : 1) * 8));
  BRA_CUDA(cudaMemcpyAsync(ctx->scratch.p, vals, (size_t)kk * 8, cudaMemcpyDefault, ctx->stream));
  BRA_CUDA(cudaMemcpyAsync(ctx->scratch2.p, V, (size_t)ld * kk * 8, cudaMemcpyDefault, ctx->stream));
  int rc = bra_pheigorth(ctx, ctx->scratch.as<double>(), ctx->scratch2.as<double>(), ld, rows, kk, orthtol);
  if (rc) return rc;
  BRA_CUDA(cudaMemcpyAsync(V, ctx->scratch2.p, (size_t)ld * kk * 8, cudaMemcpyDefault, ctx->stream));
  BRA_CUDA(cudaStreamSynchronize(ctx->stream));
  return BRA_OK;
}

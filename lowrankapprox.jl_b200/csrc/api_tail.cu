// Fused pqrfact / psvdfact drivers over the idfact core (reference: src/pqr.jl:290-307, src/psvd.jl:238-272).
#include "common.cuh"
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

int bra_set_identity(bra_ctx* ctx, int k, double* J, int64_t ldj);
int bra_chol_status_reset(bra_ctx* ctx);
int bra_lane_fork(bra_ctx* ctx);
int bra_lane_end(bra_ctx* ctx);
int bra_lane_join(bra_ctx* ctx);
int bra_chol_status(bra_ctx* ctx);
int bra_fix_signs(bra_ctx* ctx, int64_t rows, int k, double* Q, int64_t ldq, double* R, int64_t ldr);
int bra_gather_scale_cols(bra_ctx* ctx, const double* X, int64_t ldx, int64_t rows, int kk, const int* order_dev,
                          const double* scale_dev, double* out, int64_t ldo);
int bra_scatter_cols(bra_ctx* ctx, const double* src, int64_t lds, int64_t rows, int64_t n, const int64_t* jpvt1,
                     double* dst, int64_t ldd);

namespace {

struct GemmTagGuard {
  bra_ctx* c;
  explicit GemmTagGuard(bra_ctx* ctx) : c(ctx) { c->gemm_tag = BRA_PROF_TAILGEMM; }
  ~GemmTagGuard() { c->gemm_tag = BRA_PROF_GEMM; }
};

inline int64_t even(int64_t x) { return (x + 1) & ~int64_t(1); }

// After maxdet swapped columns, the sketch's R11 is no longer the triangular factor of the skeleton's sketch.  A fresh
// (k + 8)-row Gaussian sketch of C = A[:, sk] and its UNPIVOTED Householder QR (the persistent QRCP kernel with the
// pivot search switched off) give a new one in the skeleton's own column order: Omega_c C = Q_c R', so C R'^{-1} is
// well conditioned.  On return ctx->R11 holds R'; ctx->jpvt (the factorization's p) is preserved.
int fresh_preconditioner(bra_ctx* ctx, int64_t mA, int64_t k, const bra_opts* o) {
  const int64_t ldq = even(mA), nfull = ctx->res.n;
  const int64_t lc = (k + 8 < mA) ? k + 8 : mA;
  const int64_t ldt = even(mA);
  BRA_CUDA(ctx->omega_t.reserve((size_t)lc * ldt * 8));
  int rc = bra_fill_randn(ctx, ctx->omega_t.as<double>(), lc * ldt, o->seed ^ 0x6d617864657463ULL, 99);
  if (rc) return rc;
  BRA_CUDA(ctx->B2.reserve((size_t)lc * k * 8));
  double* Bc = ctx->B2.as<double>();
  if ((rc = bra_gemm_sketch(ctx, ctx->omega_t.as<double>(), lc, mA, ctx->Q.as<double>(), ldq, k, Bc, lc))) return rc;
  if ((rc = bra_allreduce_sum_f64(ctx, Bc, lc * k))) return rc;
  // the QR driver writes its (identity) pivots into ctx->jpvt: park the factorization's p meanwhile
  BRA_CUDA(ctx->aux_in2.reserve((size_t)nfull * 8));
  BRA_CUDA(cudaMemcpyAsync(ctx->aux_in2.p, ctx->jpvt.p, (size_t)nfull * 8, cudaMemcpyDeviceToDevice, ctx->stream));
  QrcpOut q = {0, 0, 0, 0};
  if ((rc = bra_qrcp_run(ctx, Bc, lc, (int)lc, k, (int)k, (int)o->nb, 0.0, 0.0, &q, /*nopivot=*/true))) return rc;
  BRA_CUDA(ctx->R11.reserve((size_t)k * k * 8));
  rc = bra_gather_R(ctx, Bc, lc, k, (int)k, ctx->jpvt.as<int64_t>(), ctx->R11.as<double>(), ctx->R11.as<double>(), k);   // R'
  BRA_CUDA(cudaMemcpyAsync(ctx->jpvt.p, ctx->aux_in2.p, (size_t)nfull * 8, cudaMemcpyDeviceToDevice, ctx->stream));
  return rc;
}

// QR of the skeleton columns: ctx->Q (mA x k, ld = even(mA)) and ctx->R1 (k x k, upper triangular).
// The sketch's R11 preconditions C = A[:, sk] only when the sketch embeds range(C): true for the oversampled Gaussian and
// SRFT sketches, NOT for :sub (a row subset: rows of low leverage are missed) and :sprn (no oversampling); there, and
// after a maxdet swap, a fresh (k + 8)-row Gaussian sketch of C is factored instead (fresh_preconditioner).  The callers
// also retry with fresh = true when the Gram Cholesky of the R11 path breaks down (the reference's Householder qr!
// cannot fail, src/pqr.jl:297-305).
bool skeleton_needs_fresh(const bra_ctx* ctx, const bra_opts* o) {
  if (ctx->res.maxdet_done) return true;
  static const bool off = getenv("BRA_SKELETON_NOFRESH") != nullptr;      // test hook: exercise the retry path
  return !off && (o->sketch == BRA_SKETCH_SUB || o->sketch == BRA_SKETCH_SPRN);
}

// defer_q (psvdfact): the second CholeskyQR pass stops after its Cholesky and the sign normalisation is skipped (U S Vt
// does not depend on it): ctx->Q holds Y1 with Q = Y1 R_y2^{-1}, R_y2 in ctx->scratch (k x k); finish_skeleton_q completes Q.
int skeleton_qr(bra_ctx* ctx, char trans, const double* dA, int64_t lda, int64_t mA, int64_t k, const bra_opts* o,
                bool fresh, bool defer_q = false) {
  const int64_t ldq = even(mA);
  BRA_CUDA(ctx->Q.reserve((size_t)ldq * k * 8));
  BRA_CUDA(ctx->R1.reserve((size_t)k * k * 8));
  int rc = bra_chol_status_reset(ctx);
  if (rc) return rc;
  rc = bra_gather_cols(ctx, trans, dA, lda, mA, k, ctx->jpvt.as<int64_t>(), ctx->Q.as<double>(), ldq);   // getcols
  if (rc) return rc;
  if (fresh && (rc = fresh_preconditioner(ctx, mA, k, o))) return rc;
  // Y = C R11^{-1}: R11 is the triangular factor of the SKETCH of these very columns (Omega*C = Q_B*R11)
  rc = bra_trsolve_right_upper(ctx, mA, (int)k, ctx->R11.as<double>(), k, ctx->Q.as<double>(), ldq);
  if (rc) return rc;
  rc = bra_cholqr2(ctx, mA, (int)k, ctx->Q.as<double>(), ldq, ctx->R11.as<double>(), ctx->R1.as<double>(), true, defer_q);
  if (rc || defer_q) return rc;
  // R1 = R_y2 R_y1 R11 inherits the signs of diag(R11) (Householder: -sign(alpha)); normalise to diag(R1) >= 0
  return bra_fix_signs(ctx, mA, (int)k, ctx->Q.as<double>(), ldq, ctx->R1.as<double>(), k);
}

// Q = Y1 R_y2^{-1} after a deferred skeleton QR, then Q' (k x mA, K-major for the U product) into ctx->Qt_l1
int finish_skeleton_q(bra_ctx* ctx, int64_t mA, int64_t k) {
  const int64_t ldq = even(mA), ldk = even(k);
  int rc = bra_trsolve_right_upper(ctx, mA, (int)k, ctx->scratch.as<double>(), k, ctx->Q.as<double>(), ldq);
  if (rc) return rc;
  return bra_transpose(ctx, ctx->Q.as<double>(), ldq, mA, k, ctx->Qt_l1.as<double>(), ldk);
}

}  // namespace

extern "C" {

int bra_pqrfact_f64(bra_ctx* ctx, char trans, int64_t m, int64_t n, const double* A, int64_t lda,
                    const bra_opts* opts, const bra_rand* rnd) {
  if (!ctx) return -1;
  int rc = bra_check_fact_args(ctx, trans, m, n, A, lda, opts);
  if (rc) return rc;
  BRA_CUDA(cudaSetDevice(ctx->device));
  const double* dA;
  int64_t dlda;
  if ((rc = bra_stage_A(ctx, trans, m, n, A, lda, opts, rnd, &dA, &dlda))) return rc;
  rc = bra_sketchfact_core(ctx, trans, m, n, dA, dlda, opts, rnd);      // V = idfact(trans, A, opts)
  if (rc) return rc;
  FactResult& res = ctx->res;
  const int64_t k = res.k, mA = res.m, nA = res.n;
  GemmTagGuard gtag(ctx);
  res.have_Q = true;
  res.have_R = true;
  if (k == 0) return bra_chol_status(ctx);
  bool fresh = skeleton_needs_fresh(ctx, opts);
  for (;;) {
    rc = skeleton_qr(ctx, trans, dA, dlda, mA, k, opts, fresh);         // F = qr!(getcols(trans, A, V[:sk]))
    if (rc) return rc;
    // R = pqrr(F.R, V[:T]) = [R1 | R1*T]   (src/pqr.jl:330-340)
    BRA_CUDA(ctx->Rfull.reserve((size_t)k * nA * 8));
    BRA_CUDA(cudaMemcpyAsync(ctx->Rfull.p, ctx->R1.p, (size_t)k * k * 8, cudaMemcpyDeviceToDevice, ctx->stream));
    if (nA > k) {
      BRA_CUDA(ctx->scratch2.reserve((size_t)even(k) * k * 8));
      rc = bra_transpose(ctx, ctx->R1.as<double>(), k, k, k, ctx->scratch2.as<double>(), even(k));   // R1^T, K-major
      if (rc) return rc;
      rc = bra_gemm_tn(ctx, ctx->scratch2.as<double>(), even(k), k, k, ctx->T.as<double>(), res.ldT, nA - k,
                       ctx->Rfull.as<double>() + (size_t)k * k, k);
      if (rc) return rc;
    }
    rc = bra_chol_status(ctx);      // the one host sync of the tail (also reports a Cholesky breakdown)
    if (rc != BRA_ERR_INTERNAL || fresh) return rc;
    fresh = true;                   // R11 was a poor preconditioner for these columns: once more with a fresh one
    ctx->skeleton_retries++;
  }
}

// prange(trans, A, opts) (src/prange.jl:14-62): an orthonormal basis Q of the range of A (trans = 'n'), of A'
// ('c'), or of both at once ('b', square A).  Result: Q, res.m x res.k (BRA_F_Q).
//   sketch = :none -> pqrfact(op(A))[:Q];  :sub -> prange_sub (:64-77) = the Q of pqrfact with the :sub sketch;
//   otherwise sketchfact(:right, trans, A, opts)[:Q]: the pivoted QR of the tall sketch B = op(A) S.
// The Q returned is the Cholesky-QR2 factor of the selected columns (diag(R) > 0): equal to the reference's
// Householder Q up to the sign of each column.
namespace {
// two_sided: one half of prange(:b, ...) -- sketchfact(:right, ...) with retval "qr", of which only the selected columns
// Q R1 = B[:, p[1:k]] are used (written to `cols`, ld = ldc); otherwise Q itself is formed (ctx->Q).
int prange_one_sided(bra_ctx* ctx, char trans, int64_t m, int64_t n, const double* dA, int64_t dlda, const bra_opts* opts,
                     const bra_rand* rnd, bool two_sided, double* cols, int64_t ldc) {
  int rc;
  const bool direct = opts->sketch == BRA_SKETCH_NONE || (opts->sketch == BRA_SKETCH_SUB && !two_sided);
  bra_opts o = *opts;
  // retval "q" alone: the swaps do not reach Q (src/pqr.jl:469-470) -- except in prange_sub (src/prange.jl:64-77), which
  // takes the COLUMNS A[:, p[1:k]] of the left sketch factorization: there the swapped p selects other columns
  if (!two_sided && opts->sketch != BRA_SKETCH_SUB) o.maxdet_tol = -1.0;
  if (direct) {
    if ((rc = bra_sketchfact_core(ctx, trans, m, n, dA, dlda, &o, rnd))) return rc;
  } else if ((rc = bra_prange_core(ctx, trans, m, n, dA, dlda, &o, rnd, two_sided))) {
    return rc;
  }
  const int64_t k = ctx->res.k, M = ctx->res.m;
  if (k == 0) return BRA_OK;
  // C = B[:, p[1:k]]: for the sketched forms rows p[1:k] of ctx->B (order x M) transposed = getcols(:c, B', p[1:k])
  const char gt = direct ? trans : 'c';
  const double* src = direct ? dA : ctx->B.as<double>();
  const int64_t lds = direct ? dlda : ctx->res.n;
  if (two_sided) return bra_gather_cols(ctx, gt, src, lds, M, k, ctx->jpvt.as<int64_t>(), cols, ldc);
  bool fresh = skeleton_needs_fresh(ctx, &o);
  for (;;) {
    if ((rc = skeleton_qr(ctx, gt, src, lds, M, k, &o, fresh))) return rc;
    rc = bra_chol_status(ctx);
    if (rc != BRA_ERR_INTERNAL || fresh) return rc;
    fresh = true;
    ctx->skeleton_retries++;
  }
}
}  // namespace

int bra_prange_f64(bra_ctx* ctx, char trans, int64_t m, int64_t n, const double* A, int64_t lda, const bra_opts* opts,
                   const bra_rand* rnd, const bra_rand* rnd2) {
  if (!ctx) return -1;
  BRA_CHECK_ARG(trans == 'n' || trans == 'c' || trans == 'b', 2, "trans");      // prange_chktrans, src/prange.jl:79-80
  int rc = bra_check_fact_args(ctx, trans == 'b' ? 'n' : trans, m, n, A, lda, opts);
  if (rc) return rc;
  if (ctx->world > 1) {
    ctx->set_error("prange on a row-sharded matrix is not built");
    return BRA_ERR_UNSUPPORTED;
  }
  if (trans == 'b' && m != n) {                                                 // checksquare, src/prange.jl:25
    ctx->set_error("DimensionMismatch: matrix is not square");
    return -3;
  }
  BRA_CUDA(cudaSetDevice(ctx->device));
  const double* dA;
  int64_t dlda;
  bra_opts onone = *opts;            // staging must not start the speculative left-sketch pipeline
  onone.sketch = BRA_SKETCH_NONE;
  if ((rc = bra_stage_A(ctx, 'n', m, n, A, lda, &onone, rnd, &dA, &dlda))) return rc;
  GemmTagGuard gtag(ctx);
  if (trans == 'b') {
    int sym = 0;
    if ((rc = bra_is_symmetric_dev(ctx, n, dA, dlda, &sym))) return rc;
    if (sym) trans = 'n';                                                        // src/prange.jl:26
  }
  if (trans != 'b') {
    if ((rc = prange_one_sided(ctx, trans, m, n, dA, dlda, opts, rnd, false, nullptr, 0))) return rc;
    ctx->res.have_Q = true;
    ctx->res.have_T = false;
    return ctx->res.k > 0 ? bra_chol_status(ctx) : BRA_OK;
  }
  // trans = :b (src/prange.jl:24-46): B = [Q_r R_r1, Q_c R_c1] = the selected (after maxdet: swapped) columns of the
  // two sketches side by side, then Q = pqrfact_backend!(B)[:Q]
  const int64_t ldc = even(n);
  BRA_CUDA(ctx->Bcat.reserve((size_t)ldc * (size_t)(2 * n > 0 ? 2 * n : 1) * 8));      // k <= n columns per side
  int64_t kk[2] = {0, 0};
  for (int side = 0; side < 2; ++side) {
    double* dst = ctx->Bcat.as<double>() + (size_t)ldc * (side == 0 ? 0 : kk[0]);
    rc = prange_one_sided(ctx, side == 0 ? 'c' : 'n', m, n, dA, dlda, opts, side == 0 ? rnd : rnd2, true, dst, ldc);
    if (rc) return rc;
    kk[side] = ctx->res.k;
  }
  onone.maxdet_tol = -1.0;           // retval "q" alone: the swaps do not reach Q (src/pqr.jl:469-470)
  const int64_t kc = kk[0] + kk[1];
  if ((rc = bra_sketchfact_core(ctx, 'n', n, kc, ctx->Bcat.as<double>(), ldc, &onone, nullptr))) return rc;
  ctx->res.have_Q = true;
  ctx->res.have_T = false;
  if (ctx->res.k == 0) return BRA_OK;
  // sketch = :none: R11 is the triangular factor of these very columns
  if ((rc = skeleton_qr(ctx, 'n', ctx->Bcat.as<double>(), ldc, n, ctx->res.k, &onone, false))) return rc;
  rc = bra_chol_status(ctx);
  if (rc != BRA_ERR_INTERNAL) return rc;
  ctx->skeleton_retries++;
  if ((rc = skeleton_qr(ctx, 'n', ctx->Bcat.as<double>(), ldc, n, ctx->res.k, &onone, true))) return rc;
  return bra_chol_status(ctx);
}

// sketchfact(side, trans, A, opts) (src/sketch.jl:52-66): the pivoted QR of the SKETCH of op(A) itself.
//   side 'l': B = S op(A) (order x n_op), its early-terminating QRCP and the ID post-processing -- exactly the first
//             stage of idfact / pqrfact / psvdfact.  Fetch BRA_F_P, BRA_F_T, BRA_F_TAU, BRA_F_BSKETCH (R on and above the
//             diagonal of the first k rows, reflectors below: Q = orgqr, R = triu(B[1:k, :])).
//   side 'r': B = op(A) S (m_op x order), the range-finder form behind prange: additionally Q (BRA_F_Q, m_op x k, the
//             CholeskyQR2 factor of B[:, p[1:k]]: the Householder Q up to the sign of each column).
int bra_sketchfact_f64(bra_ctx* ctx, char side, char trans, int64_t m, int64_t n, const double* A, int64_t lda,
                       const bra_opts* opts, const bra_rand* rnd) {
  if (!ctx) return -1;
  BRA_CHECK_ARG(side == 'l' || side == 'r', 2, "side");                 // sketchfact_chkargs, src/sketch.jl:80-84
  int rc = bra_check_fact_args(ctx, trans, m, n, A, lda, opts);
  if (rc) return rc < -1 ? rc - 1 : rc;                                  // argument numbers shift by one (side)
  if (opts->sketch == BRA_SKETCH_NONE) {
    ctx->set_error("sketchfact: sketch = :none is not a sketch (src/sketch.jl:62)");
    return -8;
  }
  BRA_CUDA(cudaSetDevice(ctx->device));
  const double* dA;
  int64_t dlda;
  if (side == 'l') {
    if ((rc = bra_stage_A(ctx, trans, m, n, A, lda, opts, rnd, &dA, &dlda))) return rc;
    if ((rc = bra_sketchfact_core(ctx, trans, m, n, dA, dlda, opts, rnd))) return rc;
    BRA_CUDA(cudaStreamSynchronize(ctx->stream));
    return BRA_OK;
  }
  if (ctx->world > 1) {
    ctx->set_error("sketchfact(:right) on a row-sharded matrix is not built");
    return BRA_ERR_UNSUPPORTED;
  }
  bra_opts onone = *opts;            // staging must not start the speculative left-sketch pipeline
  onone.sketch = BRA_SKETCH_NONE;
  if ((rc = bra_stage_A(ctx, 'n', m, n, A, lda, &onone, rnd, &dA, &dlda))) return rc;
  GemmTagGuard gtag(ctx);
  if ((rc = bra_prange_core(ctx, trans, m, n, dA, dlda, opts, rnd, true))) return rc;
  ctx->res.have_T = opts->maxdet_tol >= 0 && ctx->res.k > 0 && ctx->res.k < ctx->res.n;
  const int64_t k = ctx->res.k, M = ctx->res.m;
  if (k == 0) return BRA_OK;
  bool fresh = ctx->res.maxdet_done;
  for (;;) {
    if ((rc = skeleton_qr(ctx, 'c', ctx->B.as<double>(), ctx->res.n, M, k, opts, fresh))) return rc;
    ctx->res.have_Q = true;
    rc = bra_chol_status(ctx);
    if (rc != BRA_ERR_INTERNAL || fresh) return rc;
    fresh = true;
    ctx->skeleton_retries++;
  }
}

static int psvd_impl(bra_ctx* ctx, int64_t m, int64_t n, const double* A, int64_t lda, const bra_opts* opts,
                     const bra_rand* rnd, bool vals_only);

// Host destinations for the factors of the NEXT bra_psvdfact_f64 on this context (one shot).  U: column-major, leading
// dimension ldu >= m, room for ucols columns; S: room for scap values; Vt: leading dimension ldvt (>= the rank found).
// A factor that fits is copied to the host inside the call -- on the stream that produced it, so the copy of the
// factor that is ready first overlaps the rest of the computation -- and bra_psvd_outputs_done reports which ones were
// written (bit 0: U, 1: S, 2: Vt); the others are fetched with bra_fetch as usual.  Only page-locked (pinned)
// destinations are used.
int bra_psvd_set_outputs(bra_ctx* ctx, double* U, int64_t ldu, int64_t ucols, double* S, int64_t scap, double* Vt,
                         int64_t ldvt) {
  if (!ctx) return -1;
  // only page-locked destinations: a copy into pageable memory blocks the calling thread until it is done, which would
  // stall the enqueueing of the rest of the factorization -- such buffers are left to bra_fetch
  auto pinned = [](const void* p) {
    if (!p) return false;
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
      cudaGetLastError();
      return false;
    }
    return a.type == cudaMemoryTypeHost;
  };
  ctx->out_U = pinned(U) ? U : nullptr;
  ctx->out_ldu = ldu;
  ctx->out_ucols = ucols;
  ctx->out_S = pinned(S) ? S : nullptr;
  ctx->out_scap = scap;
  ctx->out_Vt = pinned(Vt) ? Vt : nullptr;
  ctx->out_ldvt = ldvt;
  return BRA_OK;
}
int bra_psvd_outputs_done(bra_ctx* ctx) { return ctx ? ctx->out_done : -1; }

int bra_psvdfact_f64(bra_ctx* ctx, int64_t m, int64_t n, const double* A, int64_t lda, const bra_opts* opts,
                     const bra_rand* rnd) {
  return psvd_impl(ctx, m, n, A, lda, opts, rnd, false);
}

// psvdvals(A, opts) (src/psvd.jl:274-290): the same idfact, skeleton QR and core SVD, but neither Q nor the singular
// vectors are formed -- only BRA_F_S (and bra_get_info().ksvd) are valid afterwards.
int bra_psvdvals_f64(bra_ctx* ctx, int64_t m, int64_t n, const double* A, int64_t lda, const bra_opts* opts,
                     const bra_rand* rnd) {
  return psvd_impl(ctx, m, n, A, lda, opts, rnd, true);
}

static int psvd_impl(bra_ctx* ctx, int64_t m, int64_t n, const double* A, int64_t lda, const bra_opts* opts,
                     const bra_rand* rnd, bool vals_only) {
  if (!ctx) return -1;
  ctx->out_done = 0;
  struct OutGuard {                       // registered host outputs are for this call only, however it ends
    bra_ctx* c;
    ~OutGuard() { c->out_U = c->out_S = c->out_Vt = nullptr; }
  } out_guard{ctx};
  const char trans = (m >= n) ? 'n' : 'c';                              // src/psvd.jl:242,256
  int rc = bra_check_fact_args(ctx, trans, m, n, A, lda, opts);
  if (rc) return rc;
  BRA_CUDA(cudaSetDevice(ctx->device));
  const double* dA;
  int64_t dlda;
  if ((rc = bra_stage_A(ctx, trans, m, n, A, lda, opts, rnd, &dA, &dlda))) return rc;
  rc = bra_sketchfact_core(ctx, trans, m, n, dA, dlda, opts, rnd);      // V = idfact(trans, A, opts)
  if (rc) return rc;
  FactResult& res = ctx->res;
  const int64_t k = res.k, mA = res.m, nA = res.n;
  GemmTagGuard gtag(ctx);
  res.ksvd = 0;
  res.svd_m = m;
  res.svd_n = n;
  res.svd_vals_only = vals_only;
  // copy a finished factor to its registered host destination on the CURRENT stream (main or side lane)
  auto emit_U = [&](int64_t kk_) -> int {
    if (ctx->out_U && kk_ <= ctx->out_ucols && ctx->out_ldu >= m) {
      BRA_CUDA(cudaMemcpy2DAsync(ctx->out_U, (size_t)ctx->out_ldu * 8, ctx->U.p, (size_t)m * 8, (size_t)m * 8, (size_t)kk_,
                                 cudaMemcpyDeviceToHost, ctx->stream));
      ctx->out_done |= 1;
    }
    return BRA_OK;
  };
  auto emit_Vt = [&](int64_t kk_) -> int {
    if (ctx->out_Vt && kk_ <= ctx->out_ldvt) {
      BRA_CUDA(cudaMemcpy2DAsync(ctx->out_Vt, (size_t)ctx->out_ldvt * 8, ctx->Vt.p, (size_t)kk_ * 8, (size_t)kk_ * 8, (size_t)n,
                                 cudaMemcpyDeviceToHost, ctx->stream));
      ctx->out_done |= 4;
    }
    return BRA_OK;
  };
  if (k == 0) {
    res.have_svd = true;
    return BRA_OK;
  }
  // Z = [I; T'] (nA x k): Q_z R_z by Cholesky QR  (W = R1 [I T] P' = (R1 R_z') Q_z' P').
  // Z is well conditioned (kappa(Z) = sqrt(1 + ||T||^2) / sqrt(1 + smin(T)^2), a few tens for a rank-revealing T), so
  // ONE Cholesky pass on the Gram matrix gives R_z with Q_z = Z R_z^{-1} orthonormal to eps*kappa^2, and Q_z itself is
  // never formed: Vop' = Ysel' Q_z' = (R_z^{-1} Ysel)' [I T].  A badly conditioned Z (diag(R_z) spread > 1e3) takes the
  // two-pass CholeskyQR2 with an explicit Q_z instead.
  // This chain (Z, Z'Z, R_z, R_z^{-1}) needs only T: it runs on the side lane NEXT TO the QR of the skeleton columns,
  // whose chain of small Cholesky / inverse kernels leaves most of the machine idle.
  const int64_t ldz = even(nA);
  const int64_t ldj = even(k);                // even leading dimension: 16-byte aligned columns for the Jacobi panels
  BRA_CUDA(ctx->Z.reserve((size_t)ldz * k * 8));
  BRA_CUDA(ctx->W.reserve((size_t)10 * ldj * k * 8 + 64));
  BRA_CUDA(ctx->G_l1.reserve((size_t)k * k * 8));
  BRA_CUDA(ctx->info.reserve(64));
  double* Z = ctx->Z.as<double>();
  double* Rz = ctx->W.as<double>();
  double* X = Rz + (size_t)ldj * k;         // Jacobi matrix: X = M' = R_z R1'
  double* J = X + (size_t)ldj * k;
  double* Ysel = J + (size_t)ldj * k;
  double* R2 = Ysel + (size_t)ldj * k;      // Jacobi preconditioner: X = Q2 R2 (one Cholesky pass)
  double* Q2 = R2 + (size_t)ldj * k;
  double* Xp = Q2 + (size_t)ldj * k;        // R2' (the matrix the Jacobi runs on when preconditioned)
  double* Rinv2 = Xp + (size_t)ldj * k;
  double* Rzinv = Rinv2 + (size_t)ldj * k;  // R_z^{-1} (side lane)
  double* YselU = Rzinv + (size_t)ldj * k;  // selected left vectors of the core (U chain, side lane)
  if ((size_t)k * 8 > BRA_HPIN_BYTES) {
    ctx->set_error("psvdfact: k too large for the pinned read-back buffer");
    return BRA_ERR_UNSUPPORTED;
  }
  {
    if ((rc = bra_lane_fork(ctx))) return rc;
    struct LaneEnd {
      bra_ctx* c;
      ~LaneEnd() { bra_lane_end(c); }
    } lane_end{ctx};
    if ((rc = bra_chol_status_reset(ctx))) return rc;                                       // info[13] on this lane
    if ((rc = bra_set_identity(ctx, (int)k, Z, ldz))) return rc;
    if (nA > k && (rc = bra_transpose(ctx, ctx->T.as<double>(), res.ldT, k, nA - k, Z + k, ldz))) return rc;
    if ((rc = bra_gemm_tn(ctx, Z, ldz, k, nA, Z, ldz, k, ctx->G_l1.as<double>(), k))) return rc;   // G = Z'Z = I + T T'
    if ((rc = bra_cholesky_upper(ctx, (int)k, ctx->G_l1.as<double>(), k, Rz, k))) return rc;
    if ((rc = bra_set_identity(ctx, (int)k, Rzinv, ldj))) return rc;
    if ((rc = bra_tri_inverse_upper(ctx, (int)k, Rz, k, Rzinv, ldj))) return rc;
  }
  bool fresh = skeleton_needs_fresh(ctx, opts);
skeleton_again:
  rc = skeleton_qr(ctx, trans, dA, dlda, mA, k, opts, fresh, /*defer_q=*/true);      // R = qr!(getcols(...)).R; Q deferred
  if (rc) return rc;
  bool explicit_qz = false;
  {
    if ((rc = bra_lane_join(ctx))) return rc;
    double* dz = reinterpret_cast<double*>(ctx->h_pin);
    BRA_CUDA(cudaMemcpy2DAsync(dz, 8, Rz, (size_t)(k + 1) * 8, 8, (size_t)k, cudaMemcpyDeviceToHost, ctx->stream));
    // info[12]: Cholesky status of the skeleton QR (main lane), info[13]: of Z'Z (side lane)
    BRA_CUDA(cudaMemcpyAsync(ctx->h_info + 12, ctx->info.as<int>() + 12, 8, cudaMemcpyDeviceToHost, ctx->stream));
    BRA_CUDA(cudaStreamSynchronize(ctx->stream));
    // a breakdown in the skeleton QR is an error; one in Z'Z only selects the two-pass path (which checks again)
    if (ctx->h_info[12] != 0 && !fresh) {
      // R11 was a poor preconditioner for these columns: once more with a fresh Gaussian sketch of them
      fresh = true;
      ctx->skeleton_retries++;
      goto skeleton_again;
    }
    if (ctx->h_info[12] != 0) {
      ctx->set_error("CholeskyQR2 of the skeleton columns: Gram matrix not positive definite at pivot " +
                     std::to_string(ctx->h_info[12]));
      return BRA_ERR_INTERNAL;
    }
    const int hinfo = ctx->h_info[13];
    double dmin = dz[0], dmax = dz[0];
    for (int64_t i = 0; i < k; ++i) {
      dmin = std::min(dmin, dz[i]);
      dmax = std::max(dmax, dz[i]);
    }
    explicit_qz = hinfo != 0 || !(dmin > 0.0) || dmax > 1e3 * dmin;
  }
  // Q = Y1 R_y2^{-1} and Q' are not needed before the U product: they are formed on the side lane, next to the k x k
  // products, the Cholesky / inverse of the Jacobi preconditioner and the Jacobi SVD itself (63 CTAs: most SMs are idle)
  const int64_t ldk = even(k);
  BRA_CUDA(ctx->Qt_l1.reserve((size_t)ldk * std::max(mA, nA) * 8));
  BRA_CUDA(ctx->rinv.reserve((size_t)ldk * k * 8));
  BRA_CUDA(ctx->yt.reserve((size_t)2 * ldk * mA * 8));
  if (vals_only) {
    // no singular vectors: Q = Y1 R_y2^{-1} is never needed
  } else if (explicit_qz) {
    if ((rc = finish_skeleton_q(ctx, mA, k))) return rc;       // the robust Z path below reuses ctx->scratch / rinv / yt
  } else {
    if ((rc = bra_lane_fork(ctx))) return rc;
    rc = finish_skeleton_q(ctx, mA, k);
    const int rc2 = bra_lane_end(ctx);
    if (rc || rc2) return rc ? rc : rc2;
  }
  if (explicit_qz) {
    // rebuild Z (the Gram pass left it untouched) and take the robust path; its status is checked before the SVD
    if ((rc = bra_chol_status_reset(ctx))) return rc;
    rc = bra_cholqr2(ctx, nA, (int)k, Z, ldz, nullptr, Rz);
    if (rc) return rc;
    if ((rc = bra_chol_status(ctx))) return rc;
  }
  // X[i,j] = sum_t Rz[i,t] R1[j,t]
  rc = bra_gemm_generic(ctx, Rz, 1, k, ctx->R1.as<double>(), k, 1, k, k, k, X, ldj);
  if (rc) return rc;
  std::vector<double> sig((size_t)k);
  std::vector<int> order((size_t)k);
  BRA_CUDA(ctx->S.reserve((size_t)2 * k * 8));     // [0:k] column norms (unsorted), [k:2k] sorted values
  // Jacobi preconditioning (Drmac-Veselic): X = Q2 R2 and the Jacobi runs on R2' (the rows of the triangular factor),
  // whose Gram matrix R2 R2' is far more diagonally dominant than X'X: 7 sweeps instead of 9 at C2, with most pairs
  // already converged in the last three.  X' X = D A D is graded with a benign A, so ONE Cholesky pass gives R2 with
  // Q2 = X R2^{-1} orthonormal to ~1e-12 (it only enters the right singular vectors; the left ones are the normalised
  // Jacobi columns themselves).  M = X' = Y' Sigma (Q2 J')' with R2' J' = Y' Sigma.
  bool precond = k >= 64 && getenv("BRA_JACOBI_NOPRECOND") == nullptr;
  if (precond) {
    if ((rc = bra_chol_status_reset(ctx))) return rc;
    BRA_CUDA(ctx->G.reserve((size_t)k * k * 8));
    rc = bra_gemm_tn(ctx, X, ldj, k, k, X, ldj, k, ctx->G.as<double>(), k);                 // G = X'X
    if (rc) return rc;
    rc = bra_cholesky_upper(ctx, (int)k, ctx->G.as<double>(), k, R2, ldj);                   // G = R2' R2
    if (rc) return rc;
    {
      ProfScope ps(ctx, BRA_PROF_QR);
      if ((rc = bra_set_identity(ctx, (int)k, Rinv2, ldj))) return rc;
      if ((rc = bra_tri_inverse_upper(ctx, (int)k, R2, ldj, Rinv2, ldj))) return rc;
    }
    rc = bra_gemm_generic(ctx, X, 1, ldj, Rinv2, 1, ldj, k, k, k, Q2, ldj);                  // Q2 = X R2^{-1}
    if (rc) return rc;
    rc = bra_transpose(ctx, R2, ldj, k, k, Xp, ldj);                                         // Xp = R2'
    if (rc) return rc;
    BRA_CUDA(cudaMemcpyAsync(ctx->h_info + 12, ctx->info.as<int>() + 12, 4, cudaMemcpyDeviceToHost, ctx->stream));
    BRA_CUDA(cudaStreamSynchronize(ctx->stream));
    if (ctx->h_info[12] != 0) precond = false;      // X'X not numerically positive definite: plain Jacobi on X
  }
  double* Xj = precond ? Xp : X;                   // the matrix the Jacobi orthogonalises
  // preconditioned: the Jacobi starts from the triangular R2', so the accumulated rotations can be recovered afterwards
  // as J' = R2^{-T} (Y' Sigma) (row-wise backward stable, R2 = (well conditioned) D) and the kernel need not carry J
  bool noj = precond;
  rc = bra_jacobi_svd(ctx, (int)k, Xj, ldj, J, ldj, sig.data(), order.data(), &noj);
  if (rc) return rc;
  // psvdrank (src/psvd.jl:301-308) on the sorted singular values
  std::vector<double> ssort((size_t)k);
  for (int64_t i = 0; i < k; ++i) ssort[(size_t)i] = sig[(size_t)order[(size_t)i]];
  int64_t kk = k;
  const double ptol = std::max(opts->atol, opts->rtol * ssort[0]);
  for (int64_t i = 1; i < k; ++i)
    if (ssort[(size_t)i] <= ptol) {
      kk = i;
      break;
    }
  res.ksvd = kk;
  BRA_CUDA(ctx->aux_in1.reserve((size_t)k * 4));
  BRA_CUDA(ctx->S.reserve((size_t)2 * k * 8));
  double* Ssorted = ctx->S.as<double>() + k;       // ctx->S[0:k] holds the unsorted norms (bra_jacobi_svd)
  {
    // uploads from the pinned scratch (a pageable source would make these copies synchronous)
    int* ho = reinterpret_cast<int*>(ctx->h_pin);
    double* hsrt = reinterpret_cast<double*>(ctx->h_pin + (((size_t)k * 4 + 63) & ~size_t(63)));
    if ((((size_t)k * 4 + 63) & ~size_t(63)) + (size_t)k * 8 > BRA_HPIN_BYTES) {
      ctx->set_error("psvdfact: k too large for the pinned upload buffer");
      return BRA_ERR_UNSUPPORTED;
    }
    std::memcpy(ho, order.data(), (size_t)k * 4);
    std::memcpy(hsrt, ssort.data(), (size_t)k * 8);
    BRA_CUDA(cudaMemcpyAsync(ctx->aux_in1.p, ho, (size_t)k * 4, cudaMemcpyHostToDevice, ctx->stream));
    BRA_CUDA(cudaMemcpyAsync(Ssorted, hsrt, (size_t)k * 8, cudaMemcpyHostToDevice, ctx->stream));
  }
  if (vals_only) {
    BRA_CUDA(cudaMemcpyAsync(ctx->S.p, Ssorted, (size_t)kk * 8, cudaMemcpyDeviceToDevice, ctx->stream));
    res.have_svd = true;
    BRA_CUDA(cudaStreamSynchronize(ctx->stream));
    return BRA_OK;
  }

  // Ut (kk x mA) = Jsel' Q'   and   Vp (kk x nA) = Ysel' Qz'   -- both on the TMA + DMMA kernel (TN form)
  BRA_CUDA(ctx->scratch2.reserve((size_t)ldk * std::max(mA, nA) * 8));
  BRA_CUDA(ctx->scratch3.reserve((size_t)even(kk) * std::max(mA, nA) * 8 + 64));
  BRA_CUDA(ctx->Out_l1.reserve((size_t)even(kk) * std::max(mA, nA) * 8 + 64));
  BRA_CUDA(ctx->U.reserve((size_t)std::max(m, n) * kk * 8 + 64));
  BRA_CUDA(ctx->Vt.reserve((size_t)std::max(m, n) * kk * 8 + 64));
  double* Qt = ctx->scratch2.as<double>();
  double* Out = ctx->scratch3.as<double>();
  // left factor of op(A):  Uop = Q * (left singular vectors of M)[:, order[:kk]]   (mA x kk)
  //   plain: M' J = Y Sigma  =>  left vectors of M = J;   preconditioned: left vectors = normalised columns of R2' J'
  // The U chain (one tall product + a transpose) runs on the side lane -- behind the Q it needs -- while the main stream
  // forms the right factor.  YselU is its own buffer: the right factor reuses Ysel.
  if (precond) rc = bra_gather_scale_cols(ctx, Xj, ldj, k, (int)kk, ctx->aux_in1.as<int>(), ctx->S.as<double>(), YselU, ldj);
  else rc = bra_gather_scale_cols(ctx, J, ldj, k, (int)kk, ctx->aux_in1.as<int>(), nullptr, YselU, ldj);
  if (rc) return rc;
  {
    if ((rc = bra_lane_fork(ctx))) return rc;
    struct LaneEnd {
      bra_ctx* c;
      ~LaneEnd() { bra_lane_end(c); }
    } lane_end{ctx};
    double* Uop_t = ctx->Out_l1.as<double>();          // kk x mA, ld even(kk)
    rc = bra_gemm_tn(ctx, YselU, ldj, kk, k, ctx->Qt_l1.as<double>(), ldk, mA, Uop_t, even(kk));     // (kk x mA) = Jsel' Q'
    if (rc) return rc;
    if (trans == 'n') {
      rc = bra_transpose(ctx, Uop_t, even(kk), kk, mA, ctx->U.as<double>(), mA);      // U (m x kk)
      if (rc) return rc;
      if ((rc = emit_U(kk))) return rc;
    } else {
      // op(A) = A': A ~ Vop S Uop'  =>  Vt = Uop' (kk x n), ld = kk
      BRA_CUDA(cudaMemcpy2DAsync(ctx->Vt.p, (size_t)kk * 8, Uop_t, (size_t)even(kk) * 8, (size_t)kk * 8, (size_t)mA,
                                 cudaMemcpyDeviceToDevice, ctx->stream));
      if ((rc = emit_Vt(kk))) return rc;
    }
  }
  // right factor of op(A):  Vop' = Ysel' Qz' P'   with Ysel = X[:, order] / sigma
  //   plain: right vectors of M = normalised columns of X J;   preconditioned: Q2 J'[:, order]
  if (precond) {
    if (noj) {
      // J'[:, order] = R2^{-T} (Y' Sigma)[:, order]: J's buffer is free, Rinv2 still holds R2^{-1}
      rc = bra_gather_scale_cols(ctx, Xj, ldj, k, (int)kk, ctx->aux_in1.as<int>(), nullptr, J, ldj);
      if (rc) return rc;
      rc = bra_gemm_generic(ctx, Rinv2, ldj, 1, J, 1, ldj, k, kk, k, Xp, ldj);
      if (rc) return rc;
    } else {
      // Xp (= Y' Sigma) has been consumed by the left factor above: reuse it for J'[:, order]
      rc = bra_gather_scale_cols(ctx, J, ldj, k, (int)kk, ctx->aux_in1.as<int>(), nullptr, Xp, ldj);
      if (rc) return rc;
    }
    rc = bra_gemm_generic(ctx, Q2, 1, ldj, Xp, 1, ldj, k, kk, k, Ysel, ldj);                 // Ysel = Q2 J'[:, order]
    if (rc) return rc;
  } else {
    rc = bra_gather_scale_cols(ctx, X, ldj, k, (int)kk, ctx->aux_in1.as<int>(), ctx->S.as<double>(), Ysel, ldj);
    if (rc) return rc;
  }
  BRA_CUDA(ctx->B2.reserve((size_t)even(kk) * nA * 8));
  if (explicit_qz) {
    rc = bra_transpose(ctx, Z, ldz, nA, k, Qt, ldk);                                   // Qz' (k x nA)
    if (rc) return rc;
    rc = bra_gemm_tn(ctx, Ysel, ldj, kk, k, Qt, ldk, nA, ctx->B2.as<double>(), even(kk));
    if (rc) return rc;
  } else {
    // Yh = R_z^{-1} Ysel (k x kk);  Vop' = Yh' [I T]: the first k columns are Yh' itself, the rest one TN product with T
    // R_z is well conditioned: its explicit blocked inverse (side lane, above) + one k x k x kk product
    rc = bra_gemm_generic(ctx, Rzinv, 1, ldj, Ysel, 1, ldj, k, kk, k, Q2, ldj);              // Yh = R_z^{-1} Ysel
    if (rc) return rc;
    BRA_CUDA(cudaMemcpyAsync(Ysel, Q2, (size_t)ldj * kk * 8, cudaMemcpyDeviceToDevice, ctx->stream));
    rc = bra_transpose(ctx, Ysel, ldj, k, kk, ctx->B2.as<double>(), even(kk));
    if (rc) return rc;
    if (nA > k) {
      rc = bra_gemm_tn(ctx, Ysel, ldj, kk, k, ctx->T.as<double>(), res.ldT, nA - k, ctx->B2.as<double>() + (size_t)even(kk) * k,
                       even(kk));
      if (rc) return rc;
    }
  }
  // undo the pivoting: columns j -> p[j]
  rc = bra_scatter_cols(ctx, ctx->B2.as<double>(), even(kk), kk, nA, ctx->jpvt.as<int64_t>(), Out, even(kk));
  if (rc) return rc;
  if (trans == 'n') {
    BRA_CUDA(cudaMemcpy2DAsync(ctx->Vt.p, (size_t)kk * 8, Out, (size_t)even(kk) * 8, (size_t)kk * 8, (size_t)nA,
                               cudaMemcpyDeviceToDevice, ctx->stream));
    if ((rc = emit_Vt(kk))) return rc;
  } else {
    rc = bra_transpose(ctx, Out, even(kk), kk, nA, ctx->U.as<double>(), nA);        // U = Vop (m x kk), m == nA
    if (rc) return rc;
    if ((rc = emit_U(kk))) return rc;
  }
  // singular values, sorted
  BRA_CUDA(cudaMemcpyAsync(ctx->S.p, Ssorted, (size_t)kk * 8, cudaMemcpyDeviceToDevice, ctx->stream));
  if (ctx->out_S && kk <= ctx->out_scap) {
    BRA_CUDA(cudaMemcpyAsync(ctx->out_S, ctx->S.p, (size_t)kk * 8, cudaMemcpyDeviceToHost, ctx->stream));
    ctx->out_done |= 2;
  }
  res.have_svd = true;
  // bra_fetch reports U as m x ksvd and Vt as ksvd x n of the ORIGINAL A
  res.svd_m = m;
  res.svd_n = n;
  if ((rc = bra_lane_join(ctx))) return rc;
  BRA_CUDA(cudaStreamSynchronize(ctx->stream));
  return BRA_OK;
}

}  // extern "C"

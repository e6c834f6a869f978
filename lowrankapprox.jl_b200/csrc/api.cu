// C ABI of libbrapprox: context, options, stage-wise entry points, fused drivers.
// Host-side control flow mirrors the reference's drivers:
//   sketchfact_randn            src/sketch.jl:223-240   (adaptive doubling loop)
//   geqp3_adap!                 src/pqr.jl:348-359      (rank cap)
//   pqrback_postproc, maxdet_t  src/pqr.jl:420-442
//   idfact                      src/id.jl:434-447
#include "common.cuh"
#include <cmath>
#include <cstdlib>
#include <new>

namespace {

// copy a column-major matrix between any two address spaces (host/device) on the ctx stream
cudaError_t copy2d(bra_ctx* ctx, void* dst, int64_t ldd, const void* src, int64_t lds, int64_t rows, int64_t cols,
                   size_t elem = 8) {
  if (rows <= 0 || cols <= 0) return cudaSuccess;
  if (ldd == rows && lds == rows)      // contiguous on both sides: one linear copy (full PCIe rate for host sources)
    return cudaMemcpyAsync(dst, src, (size_t)rows * (size_t)cols * elem, cudaMemcpyDefault, ctx->stream);
  return cudaMemcpy2DAsync(dst, (size_t)ldd * elem, src, (size_t)lds * elem, (size_t)rows * elem, (size_t)cols,
                           cudaMemcpyDefault, ctx->stream);
}

// returns a device pointer to the matrix (staging host data into `buf`); ld_out receives its leading dimension
int to_device(bra_ctx* ctx, DevBuf& buf, const double* src, int64_t lds, int64_t rows, int64_t cols,
              const double** out, int64_t* ld_out) {
  if (is_device_ptr(src)) {
    *out = src;
    *ld_out = lds;
    return BRA_OK;
  }
  const int64_t ld = (rows + 1) & ~int64_t(1);      // even ld keeps the TMA path available
  BRA_CUDA(buf.reserve((size_t)ld * (cols > 0 ? cols : 1) * 8));
  BRA_CUDA(copy2d(ctx, buf.p, ld, src, lds, rows, cols));
  *out = buf.as<double>();
  *ld_out = ld;
  return BRA_OK;
}

int64_t default_order(const bra_opts* o, int64_t nn) {
  // sketchfact_{randn,srft,sub}_samp defaults (src/LowRankApprox.jl:109-111); sprn uses n itself (src/sketch.jl:680)
  if (o->samp_a != 0 || o->samp_b != 0) return o->samp_a * nn + o->samp_b;
  switch (o->sketch) {
    case BRA_SKETCH_SUB: return 4 * nn + 8;
    case BRA_SKETCH_SPRN: return nn;
    default: return nn + 8;
  }
}

// Omega for one round -> K-major device copy in ctx->omega_t (ldt = roundup(mA, 2)).
int prepare_omega_t(bra_ctx* ctx, const bra_opts* o, const bra_rand* rnd, int round, int64_t order, int64_t mA) {
  const int64_t ldt = (mA + 1) & ~int64_t(1);
  BRA_CUDA(ctx->omega_t.reserve((size_t)order * ldt * 8));
  if (rnd && rnd->n_rounds > 0) {
    if (round >= rnd->n_rounds || !rnd->omega || !rnd->omega[round]) {
      ctx->set_error("adaptive loop needs more rounds than Omega matrices supplied");
      return BRA_ERR_ROUNDS;
    }
    const double* dOm;
    int64_t ldo;
    int rc = to_device(ctx, ctx->omega_in, rnd->omega[round], order, order, mA, &dOm, &ldo);
    if (rc) return rc;
    ProfScope ps(ctx, BRA_PROF_OMEGA);
    return bra_transpose_omega(ctx, dOm, ldo, order, mA, ctx->omega_t.as<double>());
  }
  ProfScope ps(ctx, BRA_PROF_OMEGA);
  if (ctx->shard_m_global > 0)      // row shard: key the stream by the global row index
    return bra_fill_randn_rows(ctx, ctx->omega_t.as<double>(), ldt, order, ctx->shard_row0,
                               (ctx->shard_m_global + 1) & ~int64_t(1), o->seed, (uint64_t)round);
  return bra_fill_randn(ctx, ctx->omega_t.as<double>(), order * ldt, o->seed, (uint64_t)round);
}

// ---- Gaussian power iteration (sketch_randn_niter > 0; src/sketch.jl:140-149, 163-172; orthrows!, src/util.jl:87-97) ----
// dst[i, jpvt[j]-1] = (i < r) ? Qp[i, j] : 0    (rows beyond the numerical rank are zero, like orthrows!(thin=false)
// zeroes the rows beyond min(m, n))
__global__ void orth_scatter_kernel(const double* __restrict__ Qp, int64_t ldq, int r, int64_t l, int64_t nn,
                                    const int64_t* __restrict__ jpvt1, double* __restrict__ dst) {
  for (int64_t j = blockIdx.x; j < nn; j += gridDim.x) {
    const double* s = Qp + j * ldq;
    double* d = dst + (jpvt1[j] - 1) * l;
    for (int64_t i = threadIdx.x; i < l; i += blockDim.x) d[i] = (i < r) ? s[i] : 0.0;
  }
}

__global__ void is_symmetric_kernel(const double* __restrict__ A, int64_t lda, int64_t n, int* __restrict__ flag) {
  // flag = 1 on any A[i,j] != A[j,i]
  for (int64_t j = blockIdx.x; j < n; j += gridDim.x)
    for (int64_t i = j + 1 + threadIdx.x; i < n; i += blockDim.x)
      if (A[i + j * lda] != A[j + i * lda]) *flag = 1;
}

// Rows of Bm (l x nn, ld = l, on the device) <- an orthonormal basis of its numerical row space, zero rows after it.
// The reference runs LAPACK's LQ (gelqf + orglq); any orthonormal basis W Q (W orthogonal) gives the same pivots, R (up
// to row signs) and T downstream, so the basis is built from the pieces this library already has: the
// early-terminating QRCP of Bm (numerical rank r, pivots), its ID  Bm P = Q_s R11 [I T], and one Cholesky pass on the
// well-conditioned Z = [I T]':  rows of R_z^{-T} [I T] P' are orthonormal and span the row space.
int orthrows_dev(bra_ctx* ctx, const bra_opts* o, int64_t l, int64_t nn, double* Bm) {
  const int64_t lmin = l < nn ? l : nn;
  QrcpOut q = {0, 0, 0, 0};
  int rc = bra_qrcp_run(ctx, Bm, l, (int)l, nn, (int)lmin, (int)o->nb, 0.0, 1e-13, &q);
  if (rc) return rc;
  const int64_t r = q.k;
  if (r <= 0) {
    BRA_CUDA(cudaMemsetAsync(Bm, 0, (size_t)l * nn * 8, ctx->stream));
    return BRA_OK;
  }
  const int64_t ldr = (r + 1) & ~int64_t(1), ldz = (nn + 1) & ~int64_t(1);
  BRA_CUDA(ctx->R11.reserve((size_t)r * r * 8));
  BRA_CUDA(ctx->T.reserve((size_t)ldr * (nn - r > 0 ? nn - r : 1) * 8));
  if ((rc = bra_gather_R(ctx, Bm, l, nn, (int)r, ctx->jpvt.as<int64_t>(), ctx->R11.as<double>(), ctx->T.as<double>(), ldr)))
    return rc;
  if (nn > r && (rc = bra_trsolve_upper(ctx, (int)r, nn - r, ctx->R11.as<double>(), r, ctx->T.as<double>(), ldr))) return rc;
  BRA_CUDA(ctx->Z.reserve((size_t)ldz * r * 8));
  double* Z = ctx->Z.as<double>();
  if ((rc = bra_set_identity(ctx, (int)r, Z, ldz))) return rc;
  if (nn > r && (rc = bra_transpose(ctx, ctx->T.as<double>(), ldr, r, nn - r, Z + r, ldz))) return rc;
  BRA_CUDA(ctx->G.reserve((size_t)r * r * 8));
  BRA_CUDA(ctx->W.reserve((size_t)2 * ldr * r * 8 + 64));
  double* Rz = ctx->W.as<double>();
  double* Rinv = Rz + (size_t)ldr * r;
  if ((rc = bra_chol_status_reset(ctx))) return rc;
  if ((rc = bra_gemm_tn(ctx, Z, ldz, r, nn, Z, ldz, r, ctx->G.as<double>(), r))) return rc;        // Z'Z = I + T T'
  if ((rc = bra_cholesky_upper(ctx, (int)r, ctx->G.as<double>(), r, Rz, ldr))) return rc;
  if ((rc = bra_set_identity(ctx, (int)r, Rinv, ldr))) return rc;
  if ((rc = bra_tri_inverse_upper(ctx, (int)r, Rz, ldr, Rinv, ldr))) return rc;
  // Qp = R_z^{-T} [I T]  (r x nn, pivoted column order)
  BRA_CUDA(ctx->B2.reserve((size_t)ldr * nn * 8));
  double* Qp = ctx->B2.as<double>();
  if ((rc = bra_transpose(ctx, Rinv, ldr, r, r, Qp, ldr))) return rc;
  if (nn > r && (rc = bra_gemm_tn(ctx, Rinv, ldr, r, r, ctx->T.as<double>(), ldr, nn - r, Qp + (size_t)ldr * r, ldr))) return rc;
  orth_scatter_kernel<<<(unsigned)(nn < 148 * 8 ? nn : 148 * 8), 128, 0, ctx->stream>>>(Qp, ldr, (int)r, l, nn,
                                                                                     ctx->jpvt.as<int64_t>(), Bm);
  ctx->launches++;
  BRA_CUDA(cudaGetLastError());
  return bra_chol_status(ctx);
}

// A' (n x m, ld = roundup(n, 2)) in ctx->At, made once per factorization (At_valid): every product with op(A) = A'
// -- the (:left, :c) sketches (src/sketch.jl:101-110), the A' passes of the power iteration (:140-149, 163-172) and the
// SRFT of the rows -- then contracts along contiguous memory, i.e. runs on the TMA + DMMA kernel like the :n form.
// One read + one write of A (HBM-bound, ~0.2 ms at 8192^2) against 2 l m n flops per round.  Returns false (no error)
// when the device has no room for the copy; the callers then take the strided generic kernel.
bool transposed_A(bra_ctx* ctx, int64_t m, int64_t n, const double* dA, int64_t lda, const double** At, int64_t* ldat) {
  const int64_t ld = (n + 1) & ~int64_t(1);
  if (!ctx->At_valid) {
    if (ctx->At.reserve((size_t)ld * (m > 0 ? m : 1) * 8) != cudaSuccess) {
      cudaGetLastError();
      return false;
    }
    if (bra_transpose(ctx, dA, lda, m, n, ctx->At.as<double>(), ld)) return false;
    ctx->At_valid = true;
  }
  *At = ctx->At.as<double>();
  *ldat = ld;
  return true;
}

// out (order x N) = Om * op(A) for a device-resident Om (order x K, ld = order); t = 'n': op(A) = A, t = 'c': A'
int apply_rows(bra_ctx* ctx, char t, int64_t m, int64_t n, const double* dA, int64_t lda, const double* Om, int64_t order,
               double* out) {
  const int64_t K = (t == 'n') ? m : n, N = (t == 'n') ? n : m;
  const int64_t ldt = (K + 1) & ~int64_t(1);
  BRA_CUDA(ctx->omega_t.reserve((size_t)order * ldt * 8));
  int rc = bra_transpose_omega(ctx, Om, order, order, K, ctx->omega_t.as<double>());
  if (rc) return rc;
  if (t == 'n') return bra_gemm_sketch(ctx, ctx->omega_t.as<double>(), order, K, dA, lda, N, out, order);
  const double* At;
  int64_t ldat;
  if (K >= 256 && transposed_A(ctx, m, n, dA, lda, &At, &ldat))
    return bra_gemm_sketch(ctx, ctx->omega_t.as<double>(), order, K, At, ldat, N, out, order);
  return bra_gemm_generic(ctx, ctx->omega_t.as<double>(), ldt, 1, dA, lda, 1, order, N, K, out, order);
}

// ishermitian(A) on the device, cached per factorization in ctx->A_sym_state (1: symmetric, -1: not)
int check_symmetric(bra_ctx* ctx, int64_t m, int64_t n, const double* dA, int64_t lda) {
  if (ctx->A_sym_state != 0) return BRA_OK;
  ctx->A_sym_state = -1;
  if (m != n) return BRA_OK;
  BRA_CUDA(ctx->info.reserve(64));
  BRA_CUDA(cudaMemsetAsync(ctx->info.as<int>() + 14, 0, 4, ctx->stream));
  is_symmetric_kernel<<<(unsigned)(n < 148 * 8 ? n : 148 * 8), 256, 0, ctx->stream>>>(dA, lda, n, ctx->info.as<int>() + 14);
  ctx->launches++;
  BRA_CUDA(cudaMemcpyAsync(ctx->h_info + 14, ctx->info.as<int>() + 14, 4, cudaMemcpyDeviceToHost, ctx->stream));
  BRA_CUDA(cudaStreamSynchronize(ctx->stream));
  if (ctx->h_info[14] == 0) ctx->A_sym_state = 1;
  return BRA_OK;
}

// the loop of sketch_randn_ln / sketch_randn_lc after the first product (ctx->B holds Bp = Omega op(A))
int power_iterations(bra_ctx* ctx, char trans, int64_t m, int64_t n, const double* dA, int64_t lda, const bra_opts* o,
                     int64_t order) {
  const int64_t nA = (trans == 'n') ? n : m, mA = (trans == 'n') ? m : n;
  const char other = (trans == 'n') ? 'c' : 'n';
  int rcs = check_symmetric(ctx, m, n, dA, lda);
  if (rcs) return rcs;
  BRA_CUDA(ctx->Bq.reserve((size_t)order * mA * 8));
  double* Bp = ctx->B.as<double>();
  double* Bq = ctx->Bq.as<double>();
  for (int it = 0; it < o->sketch_randn_niter; ++it) {
    int rc = orthrows_dev(ctx, o, order, nA, Bp);
    if (rc) return rc;
    if ((rc = apply_rows(ctx, other, m, n, dA, lda, Bp, order, Bq))) return rc;               // Bq = Bp op(A)'
    if (ctx->A_sym_state == 1) {
      BRA_CUDA(cudaMemcpyAsync(Bp, Bq, (size_t)order * nA * 8, cudaMemcpyDeviceToDevice, ctx->stream));   // Bp, Bq = Bq, Bp
    } else {
      if ((rc = orthrows_dev(ctx, o, order, mA, Bq))) return rc;
      if ((rc = apply_rows(ctx, trans, m, n, dA, lda, Bq, order, Bp))) return rc;             // Bp = Bq op(A)
    }
  }
  return BRA_OK;
}

// out (order x nA, ld = order) = Omega * op(A) for the `order` rows of Omega drawn for `round` (or supplied for it),
// summed over the row shards
int randn_product(bra_ctx* ctx, char trans, int64_t m, int64_t n, const double* dA, int64_t lda, const bra_opts* o,
                  const bra_rand* rnd, int round, int64_t order, double* out) {
  const int64_t mA = (trans == 'n') ? m : n;
  const int64_t nA = (trans == 'n') ? n : m;
  int rc = prepare_omega_t(ctx, o, rnd, round, order, mA);
  if (rc) return rc;
  ctx->sketch_rows_done += order;
  const int64_t ldt = (mA + 1) & ~int64_t(1);
  if (trans == 'n') {
    if (ctx->Apanels_state == 0) {
      // tall A: repack into row panels once per factorization if the device has room for the copy
      ctx->Apanels_state = -1;
      if (bra_gemm_wants_panels(dA, lda, mA, nA)) {
        size_t fr = 0, tot = 0;
        const size_t need = (size_t)bra_panel_bytes(mA, nA);
        if (ctx->Apanels.cap >= need ||
            (cudaMemGetInfo(&fr, &tot) == cudaSuccess && fr + ctx->Apanels.cap > need + need / 8 + (size_t(2) << 30))) {
          BRA_CUDA(ctx->Apanels.reserve(need));
          ProfScope ps(ctx, BRA_PROF_OMEGA);
          if ((rc = bra_repack_panels(ctx, dA, lda, mA, nA, ctx->Apanels.as<double>()))) return rc;
          ctx->Apanels_state = 1;
        }
      }
    }
    if (ctx->Apanels_state == 1)
      rc = bra_gemm_sketch_panels(ctx, ctx->omega_t.as<double>(), order, mA, ctx->Apanels.as<double>(), nA, out, order);
    else
      rc = bra_gemm_sketch(ctx, ctx->omega_t.as<double>(), order, mA, dA, lda, nA, out, order);
  }
  else {
    // (:left, :c): B = Omega A' contracts along the rows of A -- on the transposed copy it is the :n product
    const double* At;
    int64_t ldat;
    if (mA >= 256 && transposed_A(ctx, m, n, dA, lda, &At, &ldat))
      rc = bra_gemm_sketch(ctx, ctx->omega_t.as<double>(), order, mA, At, ldat, nA, out, order);
    else
      rc = bra_gemm_generic(ctx, ctx->omega_t.as<double>(), ldt, 1, dA, lda, 1, order, nA, mA, out, order);
  }
  if (rc) return rc;
  // row-sharded A: B = sum over ranks of Omega_g * A_g  (one all-reduce of the sketch rows per round)
  return bra_allreduce_sum_f64(ctx, out, order * nA);
}

// B (order x nA) = Omega * op(A) into ctx->B (ld = order)
int sketch_randn_round(bra_ctx* ctx, char trans, int64_t m, int64_t n, const double* dA, int64_t lda,
                       const bra_opts* o, const bra_rand* rnd, int round, int64_t order) {
  const int64_t nA = (trans == 'n') ? n : m;
  BRA_CUDA(ctx->B.reserve((size_t)order * nA * 8));
  int rc = randn_product(ctx, trans, m, n, dA, lda, o, rnd, round, order, ctx->B.as<double>());
  if (rc) return rc;
  if (o->sketch_randn_niter > 0) return power_iterations(ctx, trans, m, n, dA, lda, o, order);
  return BRA_OK;
}

// Nested Gaussian sketches (fast mode).  The library's own Omega is a counter-based stream, and the rows of round t are
// DEFINED as the rows of round t-1 followed by (order - have) rows of stream t -- so the first `have` rows of this
// round's sketch are the sketch of the previous round, kept unfactored in (raw, raw_ld), and only the new rows are
// multiplied: the adaptive loop contracts A with max(order) rows of Omega in total instead of sum(order).  Each round's
// Omega is still an i.i.d. Gaussian matrix (what the reference draws, src/sketch.jl:223-240, src/util.jl:4); what is
// given up is the independence BETWEEN rounds, which no result depends on.  Caller-supplied Omegas (parity mode), power
// iterations and BRA_SKETCH_FRESH=1 take the round-by-round path.  On return ctx->B holds the round's sketch and
// ctx->Braw a copy that survives the in-place QRCP.
int sketch_randn_nested(bra_ctx* ctx, char trans, int64_t m, int64_t n, const double* dA, int64_t lda, const bra_opts* o,
                        int round, const double* raw, int64_t raw_ld, int64_t have, int64_t order) {
  const int64_t nA = (trans == 'n') ? n : m;
  const int64_t d = order - have;
  int rc;
  BRA_CUDA(ctx->B.reserve((size_t)order * nA * 8));
  if (have <= 0) {
    if ((rc = randn_product(ctx, trans, m, n, dA, lda, o, nullptr, round, order, ctx->B.as<double>()))) return rc;
  } else if (d <= 0) {
    BRA_CUDA(copy2d(ctx, ctx->B.p, order, raw, raw_ld, order, nA));       // a sampler that does not grow: rows in hand
  } else {
    BRA_CUDA(ctx->Bnew.reserve((size_t)d * nA * 8));
    if ((rc = randn_product(ctx, trans, m, n, dA, lda, o, nullptr, round, d, ctx->Bnew.as<double>()))) return rc;
    BRA_CUDA(copy2d(ctx, ctx->B.p, order, raw, raw_ld, have, nA));
    BRA_CUDA(copy2d(ctx, ctx->B.as<double>() + have, order, ctx->Bnew.p, d, d, nA));
  }
  // the previous raw copy may be the source above: cudaFree inside reserve() waits for the copies
  BRA_CUDA(ctx->Braw.reserve((size_t)order * nA * 8));
  BRA_CUDA(cudaMemcpyAsync(ctx->Braw.p, ctx->B.p, (size_t)order * nA * 8, cudaMemcpyDeviceToDevice, ctx->stream));
  return BRA_OK;
}

// a per-round random input array: device pointer to it (staging host data into `buf`)
int stage_vec(bra_ctx* ctx, DevBuf& buf, const void* src, int64_t count, const void** out) {
  if (is_device_ptr(src)) {
    *out = src;
    return BRA_OK;
  }
  BRA_CUDA(buf.reserve((size_t)(count > 0 ? count : 1) * 8));
  BRA_CUDA(cudaMemcpyAsync(buf.p, src, (size_t)count * 8, cudaMemcpyHostToDevice, ctx->stream));
  *out = buf.p;
  return BRA_OK;
}

int need_input(bra_ctx* ctx, const bra_rand* rnd, int round, const void* const* col, const char* what) {
  if (round >= rnd->n_rounds || !col || !col[round]) {
    ctx->set_error(std::string("adaptive loop needs more rounds than random inputs supplied (") + what + ")");
    return BRA_ERR_ROUNDS;
  }
  return BRA_OK;
}

// B (order x nA) = S * op(A) into ctx->B (ld = order) for any sketch kind (src/sketch.jl:35-50 dispatch)
int sketch_round(bra_ctx* ctx, char trans, int64_t m, int64_t n, const double* dA, int64_t lda, const bra_opts* o,
                 const bra_rand* rnd, int round, int64_t order) {
  if (o->sketch == BRA_SKETCH_RANDN) return sketch_randn_round(ctx, trans, m, n, dA, lda, o, rnd, round, order);
  const int64_t mA = (trans == 'n') ? m : n;
  const int64_t nA = (trans == 'n') ? n : m;
  const bool given = rnd && rnd->n_rounds > 0;
  BRA_CUDA(ctx->B.reserve((size_t)(order > 0 ? order : 1) * nA * 8));
  int rc;
  if (o->sketch == BRA_SKETCH_SUB) {
    const void* r = nullptr;
    if (given) {
      if ((rc = need_input(ctx, rnd, round, (const void* const*)rnd->r, "r"))) return rc;
      if ((rc = stage_vec(ctx, ctx->aux_in1, rnd->r[round], order, &r))) return rc;
    } else {
      BRA_CUDA(ctx->aux_in1.reserve((size_t)order * 8));
      if ((rc = bra_fill_meta(ctx, 1, ctx->aux_in1.p, order, mA, o->seed, (uint64_t)round))) return rc;
      r = ctx->aux_in1.p;
    }
    return bra_sketch_sub(ctx, trans, dA, lda, mA, nA, order, (const int64_t*)r, ctx->B.as<double>(), order);
  }
  if (o->sketch == BRA_SKETCH_SPRN) {
    const void *perm = nullptr, *sv = nullptr;
    if (given) {
      if ((rc = need_input(ctx, rnd, round, (const void* const*)rnd->perm, "perm"))) return rc;
      if ((rc = need_input(ctx, rnd, round, (const void* const*)rnd->s, "s"))) return rc;
      if ((rc = stage_vec(ctx, ctx->aux_in1, rnd->perm[round], mA, &perm))) return rc;
      if ((rc = stage_vec(ctx, ctx->aux_in2, rnd->s[round], mA, &sv))) return rc;
    } else {
      BRA_CUDA(ctx->aux_in1.reserve((size_t)mA * 8));
      BRA_CUDA(ctx->aux_in2.reserve((size_t)(mA + 1) * 8));
      if ((rc = bra_fill_meta(ctx, 2, ctx->aux_in1.p, mA, mA, o->seed, (uint64_t)round))) return rc;
      if ((rc = bra_fill_randn(ctx, ctx->aux_in2.as<double>(), mA, o->seed, (uint64_t)round))) return rc;
      perm = ctx->aux_in1.p;
      sv = ctx->aux_in2.p;
    }
    return bra_sketch_sprn(ctx, trans, dA, lda, mA, nA, order, (const int64_t*)perm, (const double*)sv,
                           ctx->B.as<double>(), order);
  }
  if (o->sketch == BRA_SKETCH_SRFT) {
    const void *d = nullptr, *idx = nullptr;
    if (given) {
      if ((rc = need_input(ctx, rnd, round, (const void* const*)rnd->d, "d"))) return rc;
      if ((rc = need_input(ctx, rnd, round, (const void* const*)rnd->idx, "idx"))) return rc;
      if ((rc = stage_vec(ctx, ctx->aux_in2, rnd->d[round], mA, &d))) return rc;
      if ((rc = stage_vec(ctx, ctx->aux_in1, rnd->idx[round], order, &idx))) return rc;
    } else {
      BRA_CUDA(ctx->aux_in1.reserve((size_t)order * 8));
      BRA_CUDA(ctx->aux_in2.reserve((size_t)mA * 8));
      if ((rc = bra_fill_meta(ctx, 0, ctx->aux_in2.p, mA, 0, o->seed, (uint64_t)round))) return rc;
      if ((rc = bra_fill_meta(ctx, 3, ctx->aux_in1.p, order, mA, o->seed, (uint64_t)round))) return rc;
      d = ctx->aux_in2.p;
      idx = ctx->aux_in1.p;
    }
    const double* Aop = dA;
    int64_t ldop = lda;
    if (trans == 'c') {
      // sequences are the rows of A: work on a transposed copy (made once per factorization, see At_valid)
      if (!transposed_A(ctx, m, n, dA, lda, &Aop, &ldop)) {
        ctx->set_error("sketch = :srft with trans = :c needs room for a transposed copy of A on the device");
        return BRA_ERR_CUDA;
      }
    }
    return bra_sketch_srft(ctx, Aop, ldop, mA, nA, order, (const double*)d, (const int64_t*)idx, ctx->B.as<double>(),
                           order);
  }
  ctx->set_error("sketch kind not built");
  return BRA_ERR_UNSUPPORTED;
}

int run_round_qrcp(bra_ctx* ctx, const bra_opts* o, int64_t order, int64_t nA, QrcpOut* q) {
  ProfScope ps(ctx, BRA_PROF_QRCP);
  const int64_t lmin = order < nA ? order : nA;
  const int64_t kcap = (o->rank < 0 || o->rank > lmin) ? lmin : o->rank;       // src/pqr.jl:350-352
  return bra_qrcp_run(ctx, ctx->cur_B, ctx->cur_ldb, (int)order, nA, (int)kcap, (int)o->nb, o->atol, o->rtol, q);
}

}  // namespace

extern "C" {

int bra_version(void) { return 100; }

void bra_opts_default(bra_opts* o) {
  std::memset(o, 0, sizeof(*o));
  o->atol = 0.0;
  o->rtol = 5 * 2.220446049250313e-16;
  o->rank = -1;
  o->nb = 32;
  o->sketch = BRA_SKETCH_RANDN;
  o->sketch_randn_niter = 0;
  o->sketchfact_adap = 1;
  o->retval_mask = BRA_RET_Q | BRA_RET_R;
  o->maxdet_tol = -1.0;
  o->maxdet_niter = -1;
  o->samp_a = 0;
  o->samp_b = 0;
  o->seed = 0;
  o->verb = 1;
  o->flags = 0;
  o->pheig_orthtol = 1.4901161193847656e-08;      // sqrt(eps(Float64)), src/LowRankApprox.jl:102
}

int bra_create(bra_ctx** out, int device) {
  if (!out) return -1;
  *out = nullptr;
  bra_ctx* ctx = new (std::nothrow) bra_ctx();
  if (!ctx) return BRA_ERR_CUDA;
  *out = ctx;       // returned even on failure so the caller can read bra_last_error
  ctx->device = device;
  BRA_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  BRA_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) {
    ctx->set_error(std::string("libbrapprox is built for sm_100a only; device is sm_") + std::to_string(prop.major) +
                   std::to_string(prop.minor));
    return BRA_ERR_CUDA;
  }
  ctx->num_sms = prop.multiProcessorCount;
  ctx->smem_optin = (int)prop.sharedMemPerBlockOptin;
  BRA_CUDA(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
  BRA_CUDA(cudaMallocHost(&ctx->h_info, 64));
  BRA_CUDA(cudaMallocHost(&ctx->h_pin, BRA_HPIN_BYTES));
  return BRA_OK;
}

int bra_destroy(bra_ctx* ctx) {
  if (!ctx) return BRA_OK;
  cudaSetDevice(ctx->device);
  if (ctx->stream) cudaStreamSynchronize(ctx->stream);
  DevBuf* bufs[] = {&ctx->A_stage, &ctx->A_wide, &ctx->Bnew, &ctx->Braw, &ctx->omega_t, &ctx->omega_in, &ctx->B, &ctx->B2, &ctx->partial, &ctx->vn1,
                    &ctx->vn2, &ctx->lpos, &ctx->fpend, &ctx->rec, &ctx->jpvt, &ctx->tau, &ctx->rdiag, &ctx->info,
                    &ctx->kbtrace, &ctx->R11, &ctx->T, &ctx->C, &ctx->Q, &ctx->R1, &ctx->Rfull, &ctx->W, &ctx->G,
                    &ctx->U, &ctx->S, &ctx->Vt, &ctx->Z, &ctx->scratch, &ctx->scratch2, &ctx->scratch3,
                    &ctx->aux_in1, &ctx->aux_in2, &ctx->jwork, &ctx->At, &ctx->rinv, &ctx->yt, &ctx->Apanels, &ctx->tritmp, &ctx->cholscr, &ctx->Bq, &ctx->omega_spec, &ctx->Bspec, &ctx->Bt, &ctx->Bcat,
                    &ctx->partial_l1, &ctx->cholscr_l1, &ctx->tritmp_l1, &ctx->G_l1, &ctx->Qt_l1, &ctx->Out_l1};
  for (DevBuf* b : bufs) b->release();
  bra_comm_destroy(ctx);
  if (ctx->h_info) cudaFreeHost(ctx->h_info);
  if (ctx->h_pin) cudaFreeHost(ctx->h_pin);
  for (cudaEvent_t e : ctx->copy_events) cudaEventDestroy(e);
  if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
  if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
  if (ctx->ev_join) cudaEventDestroy(ctx->ev_join);
  if (ctx->side_stream) cudaStreamDestroy(ctx->side_stream);
  if (ctx->stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
  return BRA_OK;
}

const char* bra_last_error(bra_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }
uint64_t bra_launch_count(bra_ctx* ctx) { return ctx ? ctx->launches : 0; }
void* bra_stream(bra_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }

int bra_sync(bra_ctx* ctx) {
  if (!ctx) return -1;
  BRA_CUDA(cudaStreamSynchronize(ctx->stream));
  return BRA_OK;
}

int bra_profile_enable(bra_ctx* ctx, int on) {
  if (!ctx) return -1;
  BRA_CUDA(cudaStreamSynchronize(ctx->stream));
  for (auto& sp : ctx->prof_spans) {
    ctx->prof_pool.push_back(sp.e0);
    ctx->prof_pool.push_back(sp.e1);
  }
  ctx->prof_spans.clear();
  for (int t = 0; t < BRA_PROF_NTAGS; ++t) {
    ctx->prof_ms[t] = 0;
    ctx->prof_calls[t] = 0;
  }
  ctx->prof_on = on != 0;
  return BRA_OK;
}

int bra_profile_read(bra_ctx* ctx, double* ms, int64_t* calls) {
  if (!ctx) return -1;
  BRA_CUDA(cudaStreamSynchronize(ctx->stream));
  for (auto& sp : ctx->prof_spans) {
    float t = 0;
    if (cudaEventElapsedTime(&t, sp.e0, sp.e1) == cudaSuccess) {
      ctx->prof_ms[sp.tag] += t;
      ctx->prof_calls[sp.tag] += 1;
    } else {
      cudaGetLastError();
    }
    ctx->prof_pool.push_back(sp.e0);
    ctx->prof_pool.push_back(sp.e1);
  }
  ctx->prof_spans.clear();
  for (int t = 0; t < BRA_PROF_NTAGS; ++t) {
    if (ms) ms[t] = ctx->prof_ms[t];
    if (calls) calls[t] = ctx->prof_calls[t];
  }
  return BRA_OK;
}

int bra_debug_qrcp_phases(bra_ctx* ctx, int32_t* out6) {
  if (!ctx || !out6) return -1;
  for (int i = 0; i < 6; ++i) out6[i] = ctx->h_info[4 + i];
  return BRA_OK;
}

int bra_debug_jacobi_sweeps(bra_ctx* ctx) { return ctx ? ctx->last_jacobi_sweeps : -1; }
int bra_debug_skeleton_retries(bra_ctx* ctx) { return ctx ? ctx->skeleton_retries : -1; }
int64_t bra_debug_maxdet_swaps(bra_ctx* ctx) { return ctx ? ctx->last_maxdet_swaps : -1; }
int bra_debug_jacobi_phases(bra_ctx* ctx, int32_t* out4) {
  if (!ctx || !out4) return -1;
  for (int i = 0; i < 8; ++i) out4[i] = ctx->jacobi_kcycles[i];
  return BRA_OK;
}

int bra_debug_qrcp_phases_all(bra_ctx* ctx, int32_t* out, int ctas) {
  if (!ctx || !out) return -1;
  BRA_CUDA(cudaMemcpy(out, ctx->scratch3.p, (size_t)ctas * 8 * 4, cudaMemcpyDeviceToHost));
  return BRA_OK;
}

int bra_debug_qrcp_trace(bra_ctx* ctx, int64_t* out, int ctas) {
  // clock64 stamps of one pivot step of the last QRCP launch (BRA_QRCP_TS_STEP): [ctas][16 warps][16 points]
  if (!ctx || !out) return -1;
  if (ctx->scratch3.cap < (size_t)160 * 8 * 4 + (size_t)ctas * 16 * 16 * 8) return BRA_ERR_NOTREADY;
  BRA_CUDA(cudaMemcpy(out, reinterpret_cast<unsigned char*>(ctx->scratch3.p) + (size_t)160 * 8 * 4,
                      (size_t)ctas * 16 * 16 * 8, cudaMemcpyDeviceToHost));
  return BRA_OK;
}

int bra_chkopts(bra_ctx* ctx, const bra_opts* o) {
  if (!ctx) return -1;
  BRA_CHECK_ARG(o != nullptr, 2, "opts is null");
  // chkopts! (src/LowRankApprox.jl:133-141): ArgumentError("atol"/"nb"/"rtol"/"sketch")
  BRA_CHECK_ARG(o->atol >= 0, 2, "atol");
  BRA_CHECK_ARG(o->nb > 0, 2, "nb");
  BRA_CHECK_ARG(o->rtol >= 0, 2, "rtol");
  BRA_CHECK_ARG(o->sketch >= BRA_SKETCH_NONE && o->sketch <= BRA_SKETCH_SUB, 2, "sketch");
  BRA_CHECK_ARG(o->pheig_orthtol >= 0, 2, "pheig_orthtol");           // src/LowRankApprox.jl:136
  return BRA_OK;
}

int bra_sketch_randn_f64(bra_ctx* ctx, char trans, int64_t m, int64_t n, const double* A, int64_t lda,
                         int64_t order, const double* Omega, int64_t ldo, double* B, int64_t ldb) {
  if (!ctx) return -1;
  BRA_CHECK_ARG(trans == 'n' || trans == 'c', 2, "trans");            // sketch_chkargs, src/sketch.jl:76-84
  BRA_CHECK_ARG(m >= 0, 3, "m");
  BRA_CHECK_ARG(n >= 0, 4, "n");
  BRA_CHECK_ARG(A != nullptr || m * n == 0, 5, "A");
  BRA_CHECK_ARG(lda >= (m > 1 ? m : 1), 6, "lda");
  BRA_CHECK_ARG(order >= 0, 7, "order");
  const int64_t mA = (trans == 'n') ? m : n, nA = (trans == 'n') ? n : m;
  BRA_CHECK_ARG(Omega != nullptr || order * mA == 0, 8, "Omega");
  BRA_CHECK_ARG(ldo >= (order > 1 ? order : 1), 9, "ldo");
  BRA_CHECK_ARG(B != nullptr || order * nA == 0, 10, "B");
  BRA_CHECK_ARG(ldb >= (order > 1 ? order : 1), 11, "ldb");
  if (order == 0 || nA == 0) return BRA_OK;
  BRA_CUDA(cudaSetDevice(ctx->device));
  const double* dA;
  int64_t dlda;
  int rc = to_device(ctx, ctx->A_stage, A, lda, m, n, &dA, &dlda);
  if (rc) return rc;
  bra_opts o;
  bra_opts_default(&o);
  const double* oms[1] = {Omega};
  bra_rand rnd;
  std::memset(&rnd, 0, sizeof(rnd));
  rnd.n_rounds = 1;
  rnd.omega = oms;
  // honour ldo by staging through to_device inside prepare_omega_t: it assumes ld == order, so compact first
  if (ldo != order) {
    BRA_CUDA(ctx->scratch2.reserve((size_t)order * mA * 8));
    BRA_CUDA(copy2d(ctx, ctx->scratch2.p, order, Omega, ldo, order, mA));
    oms[0] = ctx->scratch2.as<double>();
  }
  ctx->Apanels_state = 0;
  ctx->At_valid = false;
  rc = sketch_randn_round(ctx, trans, m, n, dA, dlda, &o, &rnd, 0, order);
  if (rc) return rc;
  BRA_CUDA(copy2d(ctx, B, ldb, ctx->B.p, order, order, nA));
  BRA_CUDA(cudaStreamSynchronize(ctx->stream));
  return BRA_OK;
}

// Stage-wise sketch for the three structured kinds: `kind` is BRA_SKETCH_{SPRN,SRFT,SUB}; v1/v2 are the random
// inputs in reference order (sub: r, -; sprn: perm, s; srft: d, idx).
static int sketch_structured(bra_ctx* ctx, int kind, char trans, int64_t m, int64_t n, const double* A, int64_t lda,
                             int64_t order, const void* v1, const void* v2, double* B, int64_t ldb) {
  BRA_CHECK_ARG(trans == 'n' || trans == 'c', 2, "trans");
  BRA_CHECK_ARG(m >= 0, 3, "m");
  BRA_CHECK_ARG(n >= 0, 4, "n");
  BRA_CHECK_ARG(A != nullptr || m * n == 0, 5, "A");
  BRA_CHECK_ARG(lda >= (m > 1 ? m : 1), 6, "lda");
  BRA_CHECK_ARG(order >= 0, 7, "order");
  const int64_t nA = (trans == 'n') ? n : m, mA = (trans == 'n') ? m : n;
  BRA_CHECK_ARG(v1 != nullptr || order * mA == 0, 8, "random input 1");
  BRA_CHECK_ARG(kind == BRA_SKETCH_SUB || v2 != nullptr || order * mA == 0, 9, "random input 2");
  BRA_CHECK_ARG(B != nullptr || order * nA == 0, 10, "B");
  BRA_CHECK_ARG(ldb >= (order > 1 ? order : 1), 11, "ldb");
  if (order == 0 || nA == 0) return BRA_OK;
  if (mA == 0) {
    BRA_CUDA(cudaMemset2DAsync(B, (size_t)ldb * 8, 0, (size_t)order * 8, (size_t)nA, ctx->stream));
    BRA_CUDA(cudaStreamSynchronize(ctx->stream));
    return BRA_OK;
  }
  BRA_CUDA(cudaSetDevice(ctx->device));
  const double* dA;
  int64_t dlda;
  int rc = to_device(ctx, ctx->A_stage, A, lda, m, n, &dA, &dlda);
  if (rc) return rc;
  bra_opts o;
  bra_opts_default(&o);
  o.sketch = kind;
  const void* a1[1] = {v1};
  const void* a2[1] = {v2};
  bra_rand rnd;
  std::memset(&rnd, 0, sizeof(rnd));
  rnd.n_rounds = 1;
  if (kind == BRA_SKETCH_SUB) {
    rnd.r = (const int64_t* const*)a1;
  } else if (kind == BRA_SKETCH_SPRN) {
    rnd.perm = (const int64_t* const*)a1;
    rnd.s = (const double* const*)a2;
  } else {
    rnd.d = (const double* const*)a1;
    rnd.idx = (const int64_t* const*)a2;
  }
  ctx->At_valid = false;
  ctx->Apanels_state = -1;
  rc = sketch_round(ctx, trans, m, n, dA, dlda, &o, &rnd, 0, order);
  if (rc) return rc;
  BRA_CUDA(copy2d(ctx, B, ldb, ctx->B.p, order, order, nA));
  BRA_CUDA(cudaStreamSynchronize(ctx->stream));
  return BRA_OK;
}

int bra_sketch_sub_f64(bra_ctx* ctx, char trans, int64_t m, int64_t n, const double* A, int64_t lda, int64_t order,
                       const int64_t* r, double* B, int64_t ldb) {
  if (!ctx) return -1;
  return sketch_structured(ctx, BRA_SKETCH_SUB, trans, m, n, A, lda, order, r, nullptr, B, ldb);
}

int bra_sketch_sprn_f64(bra_ctx* ctx, char trans, int64_t m, int64_t n, const double* A, int64_t lda, int64_t order,
                        const int64_t* perm, const double* s, double* B, int64_t ldb) {
  if (!ctx) return -1;
  return sketch_structured(ctx, BRA_SKETCH_SPRN, trans, m, n, A, lda, order, perm, s, B, ldb);
}

int bra_sketch_srft_f64(bra_ctx* ctx, char trans, int64_t m, int64_t n, const double* A, int64_t lda, int64_t order,
                        const double* d, const int64_t* idx, double* B, int64_t ldb) {
  if (!ctx) return -1;
  return sketch_structured(ctx, BRA_SKETCH_SRFT, trans, m, n, A, lda, order, d, idx, B, ldb);
}

int bra_geqp3_adap_f64(bra_ctx* ctx, int64_t l, int64_t n, double* B, int64_t ldb, const bra_opts* opts,
                       int64_t* jpvt, double* tau, int64_t* k, int64_t* nsteps, int32_t* kb_trace, int64_t kb_cap,
                       int64_t* n_blocks) {
  if (!ctx) return -1;
  BRA_CHECK_ARG(l >= 0 && l < (int64_t(1) << 30), 2, "l");
  BRA_CHECK_ARG(n >= 0, 3, "n");
  BRA_CHECK_ARG(B != nullptr || l * n == 0, 4, "B");
  BRA_CHECK_ARG(ldb >= (l > 1 ? l : 1), 5, "ldb");
  int rc = bra_chkopts(ctx, opts);
  if (rc) return -6;
  BRA_CUDA(cudaSetDevice(ctx->device));
  const int64_t lmin = l < n ? l : n;
  const int64_t kcap = (opts->rank < 0 || opts->rank > lmin) ? lmin : opts->rank;
  QrcpOut q = {0, 0, 0, 0};
  if (kcap > 0) {
    BRA_CUDA(ctx->B.reserve((size_t)l * n * 8));
    BRA_CUDA(copy2d(ctx, ctx->B.p, l, B, ldb, l, n));
    {
      ProfScope ps(ctx, BRA_PROF_QRCP);
      rc = bra_qrcp_run(ctx, ctx->B.as<double>(), l, (int)l, n, (int)kcap, (int)opts->nb, opts->atol, opts->rtol, &q);
    }
    if (rc) return rc;
    BRA_CUDA(ctx->B2.reserve((size_t)l * n * 8));
    rc = bra_permute_cols(ctx, ctx->B.as<double>(), l, ctx->B2.as<double>(), l, l, n, ctx->jpvt.as<int64_t>());
    if (rc) return rc;
    BRA_CUDA(copy2d(ctx, B, ldb, ctx->B2.p, l, l, n));
    if (jpvt) BRA_CUDA(cudaMemcpyAsync(jpvt, ctx->jpvt.p, (size_t)n * 8, cudaMemcpyDefault, ctx->stream));
    if (tau && q.nsteps > 0)
      BRA_CUDA(cudaMemcpyAsync(tau, ctx->tau.p, (size_t)q.nsteps * 8, cudaMemcpyDefault, ctx->stream));
    if (kb_trace && kb_cap > 0 && q.nblocks > 0) {
      int64_t c = q.nblocks < kb_cap ? q.nblocks : kb_cap;
      BRA_CUDA(cudaMemcpyAsync(kb_trace, ctx->kbtrace.p, (size_t)c * 4, cudaMemcpyDefault, ctx->stream));
    }
    BRA_CUDA(cudaStreamSynchronize(ctx->stream));
  } else if (jpvt) {
    std::vector<int64_t> id((size_t)n);
    for (int64_t j = 0; j < n; ++j) id[(size_t)j] = j + 1;
    BRA_CUDA(cudaMemcpy(jpvt, id.data(), (size_t)n * 8, cudaMemcpyDefault));
  }
  if (k) *k = q.k;
  if (nsteps) *nsteps = q.nsteps;
  if (n_blocks) *n_blocks = q.nblocks;
  return BRA_OK;
}

int bra_trsolve_T_f64(bra_ctx* ctx, int64_t k, int64_t n, const double* R, int64_t ldr, double* T, int64_t ldt) {
  if (!ctx) return -1;
  BRA_CHECK_ARG(k >= 0 && k < (int64_t(1) << 30), 2, "k");
  BRA_CHECK_ARG(n >= k, 3, "n");
  BRA_CHECK_ARG(R != nullptr || k * n == 0, 4, "R");
  BRA_CHECK_ARG(ldr >= (k > 1 ? k : 1), 5, "ldr");
  BRA_CHECK_ARG(T != nullptr || k * (n - k) == 0, 6, "T");
  BRA_CHECK_ARG(ldt >= (k > 1 ? k : 1), 7, "ldt");
  if (k == 0 || n == k) return BRA_OK;
  BRA_CUDA(cudaSetDevice(ctx->device));
  BRA_CUDA(ctx->R11.reserve((size_t)k * k * 8));
  BRA_CUDA(ctx->T.reserve((size_t)k * (n - k) * 8));
  BRA_CUDA(copy2d(ctx, ctx->R11.p, k, R, ldr, k, k));
  BRA_CUDA(copy2d(ctx, ctx->T.p, k, R + k * ldr, ldr, k, n - k));
  int rc = bra_trsolve_upper(ctx, (int)k, n - k, ctx->R11.as<double>(), k, ctx->T.as<double>(), k);
  if (rc) return rc;
  BRA_CUDA(copy2d(ctx, T, ldt, ctx->T.p, k, k, n - k));
  BRA_CUDA(cudaStreamSynchronize(ctx->stream));
  return BRA_OK;
}

// nested Gaussian sketches apply to the library's own Omega only (see sketch_randn_nested)
bool bra_sketch_nested(const bra_opts* o, const bra_rand* rnd) {
  return o->sketch == BRA_SKETCH_RANDN && !(rnd && rnd->n_rounds > 0) && o->sketch_randn_niter == 0 &&
         !(o->flags & BRA_OPT_FRESH_SKETCH) && getenv("BRA_SKETCH_FRESH") == nullptr;
}

// sketchfact(:left, trans, A, opts) + pqrback_postproc for retval "t"; shared by idfact/pqrfact/psvdfact.
int bra_sketchfact_core(bra_ctx* ctx, char trans, int64_t m, int64_t n, const double* dA, int64_t lda,
                        const bra_opts* o, const bra_rand* rnd) {
  const int64_t nA = (trans == 'n') ? n : m;
  // the speculative sketches of bra_stage_A belong to THIS call only: whatever way it ends (errors included), a later
  // factorization must never find them
  struct SpecGuard {
    bra_ctx* c;
    ~SpecGuard() { c->spec_rounds = 0; }
  } spec_guard{ctx};
  FactResult& res = ctx->res;
  res = FactResult();
  res.m = (trans == 'n') ? m : n;
  res.n = nA;
  ctx->At_valid = false;
  ctx->Apanels_state = 0;
  ctx->A_sym_state = 0;
  if (ctx->spec_rounds == 0) ctx->sketch_rows_done = 0;          // (else: the stacked rows of bra_stage_A)
  QrcpOut q = {0, 0, 0, 0};
  int64_t order = 0;
  if (o->sketch == BRA_SKETCH_NONE) {
    // sketch = :none (pqrfact_none, src/pqr.jl:323-327): the early-terminating QRCP runs on a copy of op(A) itself.
    // Same persistent kernel (generic row loop; the column slabs of a tall matrix live in L2/HBM, not in shared
    // memory), so this is the classical expensive path -- functional, not a benchmarked one.
    const int64_t mA = res.m;
    order = mA;
    if (mA >= (int64_t(1) << 14) + 4096 && !bra_qrcp_blocked_ok(mA, nA, (int)(o->nb < 32 ? o->nb : 32), ctx->num_sms)) {
      ctx->set_error("sketch = :none on more than ~20000 rows needs the blocked kernel: nb <= 32 and at most 128 columns "
                     "of op(A) per SM");
      return BRA_ERR_UNSUPPORTED;
    }
    BRA_CUDA(ctx->B.reserve((size_t)(mA > 0 ? mA : 1) * (nA > 0 ? nA : 1) * 8));
    int rc;
    if (trans == 'n') {
      BRA_CUDA(copy2d(ctx, ctx->B.p, mA, dA, lda, mA, nA));
    } else if ((rc = bra_transpose(ctx, dA, lda, m, n, ctx->B.as<double>(), mA))) {
      return rc;
    }
    ctx->cur_B = ctx->B.as<double>();
    ctx->cur_ldb = order;
    ctx->cur_rows = order;
    rc = run_round_qrcp(ctx, o, order, nA, &q);
    if (rc) return rc;
    res.orders[0] = order;
    res.ks[0] = q.k;
    res.steps[0] = q.nsteps;
    res.rounds = 1;
  } else if (o->sketchfact_adap || o->rank < 0) {
    int64_t nn = o->nb << ctx->start_round;                          // src/sketch.jl:226 (n doubles every round)
    const bool nested = bra_sketch_nested(o, rnd);
    const double* raw = nullptr;                                     // nested: the unfactored sketch rows so far
    int64_t raw_ld = 0, have = 0;
    for (int round = ctx->start_round;; ++round) {
      if (round >= BRA_MAX_ROUNDS) {
        ctx->set_error("adaptive loop exceeded BRA_MAX_ROUNDS");
        return BRA_ERR_ROUNDS;
      }
      order = default_order(o, nn);
      int rc = BRA_OK;
      if (round < ctx->spec_rounds && nested) {
        // formed while A was still arriving from the host (bra_stage_A): the first `order` rows of the stacked sketch,
        // copied out because the QRCP works in place and the next round needs them unfactored
        BRA_CUDA(ctx->B.reserve((size_t)order * nA * 8));
        BRA_CUDA(copy2d(ctx, ctx->B.p, order, ctx->Bspec.p, ctx->spec_ld, order, nA));
        ctx->cur_B = ctx->B.as<double>();
        ctx->cur_ldb = order;
        ctx->cur_rows = order;
        raw = ctx->Bspec.as<double>();
        raw_ld = ctx->spec_ld;
        have = order;
      } else if (round < ctx->spec_rounds) {
        // this round's sketch was formed while A was still arriving from the host (bra_stage_A)
        ctx->cur_B = ctx->Bspec.as<double>() + ctx->spec_off[round];
        ctx->cur_ldb = ctx->spec_ld;
        ctx->cur_rows = order;
      } else if (nested) {
        rc = sketch_randn_nested(ctx, trans, m, n, dA, lda, o, round, raw, raw_ld, have, order);
        ctx->cur_B = ctx->B.as<double>();
        ctx->cur_ldb = order;
        ctx->cur_rows = order;
        raw = ctx->Braw.as<double>();
        raw_ld = order;
        have = order;
      } else {
        rc = sketch_round(ctx, trans, m, n, dA, lda, o, rnd, round, order);
        ctx->cur_B = ctx->B.as<double>();
        ctx->cur_ldb = order;
        ctx->cur_rows = order;
      }
      if (rc) return rc;
      rc = run_round_qrcp(ctx, o, order, nA, &q);
      if (rc) return rc;
      res.orders[round] = order;
      res.ks[round] = q.k;
      res.steps[round] = q.nsteps;
      res.rounds = round + 1;
      if (q.k < nn) break;                                            // src/sketch.jl:232
      nn *= 2;
    }
  } else {
    order = (o->sketch == BRA_SKETCH_SPRN) ? o->rank : default_order(o, o->rank);   // src/sketch.jl:236,686
    int rc = sketch_round(ctx, trans, m, n, dA, lda, o, rnd, 0, order);
    if (rc) return rc;
    ctx->cur_B = ctx->B.as<double>();
    ctx->cur_ldb = order;
    ctx->cur_rows = order;
    rc = run_round_qrcp(ctx, o, order, nA, &q);
    if (rc) return rc;
    res.orders[0] = order;
    res.ks[0] = q.k;
    res.steps[0] = q.nsteps;
    res.rounds = 1;
  }
  res.k = q.k;
  if (q.nsteps == 0) {
    // kcap == 0: identity permutation
    BRA_CUDA(ctx->jpvt.reserve((size_t)nA * 8));
    std::vector<int64_t> id((size_t)nA);
    for (int64_t j = 0; j < nA; ++j) id[(size_t)j] = j + 1;
    BRA_CUDA(cudaMemcpyAsync(ctx->jpvt.p, id.data(), (size_t)nA * 8, cudaMemcpyHostToDevice, ctx->stream));
    BRA_CUDA(cudaStreamSynchronize(ctx->stream));
  }
  // pqrback_postproc: R = triu(B[1:k,:]); T = R11 \ R12   (src/pqr.jl:428-429, 438-442)
  const int64_t k = res.k;
  if (k > 0) {
    BRA_CUDA(ctx->R11.reserve((size_t)k * k * 8));
    const int64_t ldT = (k + 1) & ~int64_t(1);
    res.ldT = ldT;
    BRA_CUDA(ctx->T.reserve((size_t)ldT * (nA - k > 0 ? nA - k : 1) * 8));
    int rc;
    {
      ProfScope ps(ctx, BRA_PROF_GATHER);
      rc = bra_gather_R(ctx, ctx->cur_B, ctx->cur_ldb, nA, (int)k, ctx->jpvt.as<int64_t>(), ctx->R11.as<double>(),
                        ctx->T.as<double>(), ldT);
    }
    if (rc) return rc;
    {
      ProfScope ps(ctx, BRA_PROF_TRSOLVE);
      rc = bra_trsolve_upper(ctx, (int)k, nA - k, ctx->R11.as<double>(), k, ctx->T.as<double>(), ldT);
    }
    if (rc) return rc;
    // strong RRQR post-processing (pqrback_postproc, src/pqr.jl:428-433): only p and T are maintained
    ctx->last_maxdet_swaps = 0;
    if (o->maxdet_tol >= 0 && k < nA) {
      ProfScope ps(ctx, BRA_PROF_TRSOLVE);
      rc = bra_maxdet_swapcols(ctx, (int)k, nA - k, ctx->T.as<double>(), ldT, ctx->jpvt.as<int64_t>(), o->maxdet_tol,
                               o->maxdet_niter, &ctx->last_maxdet_swaps);
      if (rc) return rc;
      res.maxdet_done = ctx->last_maxdet_swaps > 0;
    }
  }
  res.have_T = true;
  ctx->spec_rounds = 0;
  return BRA_OK;
}

// sketchfact(:right, trans, A, opts) (src/sketch.jl:52-66 with side = :right; the drivers at :213-236, :313-330,
// :545-562, :674-690): B = op(A) S is the transpose of the left sketch of op(A)' on the same random inputs (the four
// mul! forms of every sketch type are transposes of each other in real arithmetic), so each round forms
// S' op(A)' (order x M) with the left-sketch kernels, transposes it to the tall M x order matrix B and runs the
// early-terminating QRCP on B's columns (the persistent kernel's tall-slab path, as for sketch = :none).
// On return: ctx->B holds the last round's S' op(A)' (= B', intact), ctx->jpvt the column pivots of B, ctx->R11 the
// triangular factor of B[:, p[1:k]], res.m = M, res.n = order, res.k = k.  with_maxdet: also T and the maxdet swaps
// (only the two-sided form keeps them: with retval "q" alone the reference's Q ignores the swaps, src/pqr.jl:469-470).
int bra_prange_core(bra_ctx* ctx, char trans, int64_t m, int64_t n, const double* dA, int64_t lda, const bra_opts* o,
                    const bra_rand* rnd, bool with_maxdet) {
  const char ft = (trans == 'n') ? 'c' : 'n';
  const int64_t M = (trans == 'n') ? m : n;                           // rows of op(A) = rows of B
  FactResult& res = ctx->res;
  res = FactResult();
  res.m = M;
  ctx->At_valid = false;
  ctx->Apanels_state = 0;
  ctx->A_sym_state = 0;
  ctx->spec_rounds = 0;
  if (M >= (int64_t(1) << 14) + 4096 && o->nb > 32) {
    ctx->set_error("prange on more than ~20000 rows of op(A) needs the blocked QR kernel: nb <= 32");
    return BRA_ERR_UNSUPPORTED;
  }
  QrcpOut q = {0, 0, 0, 0};
  int64_t order = 0;
  const bool adaptive = o->sketchfact_adap || o->rank < 0;
  int64_t nn = adaptive ? o->nb : o->rank;
  for (int round = 0;; ++round) {
    if (round >= BRA_MAX_ROUNDS) {
      ctx->set_error("adaptive loop exceeded BRA_MAX_ROUNDS");
      return BRA_ERR_ROUNDS;
    }
    order = (!adaptive && o->sketch == BRA_SKETCH_SPRN) ? o->rank : default_order(o, nn);
    int rc = sketch_round(ctx, ft, m, n, dA, lda, o, rnd, round, order);
    if (rc) return rc;
    BRA_CUDA(ctx->Bt.reserve((size_t)(M > 0 ? M : 1) * (order > 0 ? order : 1) * 8));
    if (M > 0 && order > 0 && (rc = bra_transpose(ctx, ctx->B.as<double>(), order, order, M, ctx->Bt.as<double>(), M)))
      return rc;
    ctx->cur_B = ctx->Bt.as<double>();
    ctx->cur_ldb = M > 0 ? M : 1;
    ctx->cur_rows = M;
    if ((rc = run_round_qrcp(ctx, o, M, order, &q))) return rc;
    res.orders[round] = order;
    res.ks[round] = q.k;
    res.steps[round] = q.nsteps;
    res.rounds = round + 1;
    if (!adaptive || q.k < nn) break;                                 // src/sketch.jl:232
    nn *= 2;
  }
  res.n = order;
  res.k = q.k;
  if (q.nsteps == 0) {
    BRA_CUDA(ctx->jpvt.reserve((size_t)(order > 0 ? order : 1) * 8));
    std::vector<int64_t> id((size_t)order);
    for (int64_t j = 0; j < order; ++j) id[(size_t)j] = j + 1;
    BRA_CUDA(cudaMemcpyAsync(ctx->jpvt.p, id.data(), (size_t)order * 8, cudaMemcpyHostToDevice, ctx->stream));
    BRA_CUDA(cudaStreamSynchronize(ctx->stream));
  }
  const int64_t k = res.k;
  if (k > 0) {
    BRA_CUDA(ctx->R11.reserve((size_t)k * k * 8));
    const int64_t ldT = (k + 1) & ~int64_t(1);
    res.ldT = ldT;
    BRA_CUDA(ctx->T.reserve((size_t)ldT * (order - k > 0 ? order - k : 1) * 8));
    int rc = bra_gather_R(ctx, ctx->cur_B, ctx->cur_ldb, order, (int)k, ctx->jpvt.as<int64_t>(), ctx->R11.as<double>(),
                          ctx->T.as<double>(), ldT);
    if (rc) return rc;
    ctx->last_maxdet_swaps = 0;
    if (with_maxdet && o->maxdet_tol >= 0 && k < order) {
      if ((rc = bra_trsolve_upper(ctx, (int)k, order - k, ctx->R11.as<double>(), k, ctx->T.as<double>(), ldT))) return rc;
      rc = bra_maxdet_swapcols(ctx, (int)k, order - k, ctx->T.as<double>(), ldT, ctx->jpvt.as<int64_t>(), o->maxdet_tol,
                               o->maxdet_niter, &ctx->last_maxdet_swaps);
      if (rc) return rc;
      res.maxdet_done = ctx->last_maxdet_swaps > 0;
    }
  }
  return BRA_OK;
}

int bra_is_symmetric_dev(bra_ctx* ctx, int64_t n, const double* dA, int64_t lda, int* sym) {
  ctx->A_sym_state = 0;
  int rc = check_symmetric(ctx, n, n, dA, lda);
  *sym = ctx->A_sym_state == 1;
  ctx->A_sym_state = 0;
  return rc;
}

int bra_check_fact_args(bra_ctx* ctx, char trans, int64_t m, int64_t n, const double* A, int64_t lda,
                        const bra_opts* opts) {
  BRA_CHECK_ARG(trans == 'n' || trans == 'c', 2, "trans");            // chktrans, src/LowRankApprox.jl:150
  BRA_CHECK_ARG(m >= 0, 3, "m");
  BRA_CHECK_ARG(n >= 0, 4, "n");
  BRA_CHECK_ARG(A != nullptr || m * n == 0, 5, "A");
  BRA_CHECK_ARG(lda >= (m > 1 ? m : 1), 6, "lda");
  if (bra_chkopts(ctx, opts)) return -7;
  if (opts->sketch_randn_niter > 0 && ctx->world > 1) {
    ctx->set_error("sketch_randn_niter > 0 on a row-sharded matrix is not built");
    return BRA_ERR_UNSUPPORTED;
  }
  if (opts->sketch == BRA_SKETCH_NONE && ctx->world > 1) {
    ctx->set_error("sketch = :none on a row-sharded matrix is not built");
    return BRA_ERR_UNSUPPORTED;
  }
  if (ctx->world > 1 && (opts->sketch != BRA_SKETCH_RANDN || trans != 'n')) {
    ctx->set_error("row-sharded factorizations need sketch = :randn and trans = 'n' (the sharded dimension must be "
                   "the contracted one)");
    return BRA_ERR_UNSUPPORTED;
  }
  return BRA_OK;
}

// Device copy of A for a fused factorization.  A device-resident A is used in place.  A large host-resident A headed
// for the adaptive Gaussian path is uploaded in column panels on a second stream while the sketch products of the
// first adaptive rounds run on the panels that have already arrived: all those rounds share one stacked Omega (the
// same Philox streams the round-by-round path draws), so by the time the upload ends their sketches exist and only the
// pivoted QRs remain.  Rounds beyond the speculative ones fall back to the ordinary per-round sketch; sketches of rounds
// that turn out not to be needed cost nothing visible (they hide behind the PCIe transfer).
int bra_stage_A(bra_ctx* ctx, char trans, int64_t m, int64_t n, const double* A, int64_t lda, const bra_opts* o,
                const bra_rand* rnd, const double** dA, int64_t* dlda) {
  ctx->spec_rounds = 0;
  if (is_device_ptr(A)) {
    *dA = A;
    *dlda = lda;
    return BRA_OK;
  }
  const int64_t ld = (m + 1) & ~int64_t(1);
  BRA_CUDA(ctx->A_stage.reserve((size_t)ld * (n > 0 ? n : 1) * 8));
  *dA = ctx->A_stage.as<double>();
  *dlda = ld;
  if (m <= 0 || n <= 0) return BRA_OK;
  const bool spec = trans == 'n' && o->sketch == BRA_SKETCH_RANDN && (o->sketchfact_adap || o->rank < 0) &&
                    !(rnd && rnd->n_rounds > 0) && o->sketch_randn_niter == 0 && ctx->world == 1 &&
                    ctx->start_round == 0 && (int64_t)m * n * 8 >= (int64_t(64) << 20) && n >= 2048 &&
                    getenv("BRA_NO_SPEC_UPLOAD") == nullptr;
  if (!spec) {
    BRA_CUDA(copy2d(ctx, ctx->A_stage.p, ld, A, lda, m, n));
    return BRA_OK;
  }
  // speculative rounds: orders 40, 72, 136, 264, 520 with the default sampler (stop before the order passes min(m, n))
  int T = 0;
  int64_t lsum = 0, orders[5], fresh_rows[5];
  const bool nested = bra_sketch_nested(o, rnd);        // round t = the first orders[t] rows of ONE stacked sketch
  for (int64_t nn = o->nb; T < 5; nn *= 2, ++T) {
    const int64_t ord = default_order(o, nn);
    if (ord > m || ord > n || ord <= 0 || (nested && ord <= lsum)) break;
    orders[T] = ord;
    ctx->spec_off[T] = nested ? 0 : lsum;
    fresh_rows[T] = nested ? ord - lsum : ord;          // rows drawn from stream T, stored from row (nested ? lsum : spec_off)
    lsum = nested ? ord : lsum + ord;
  }
  if (T == 0) {
    BRA_CUDA(copy2d(ctx, ctx->A_stage.p, ld, A, lda, m, n));
    return BRA_OK;
  }
  const int64_t ldt = ld;
  BRA_CUDA(ctx->omega_spec.reserve((size_t)lsum * ldt * 8));
  BRA_CUDA(ctx->Bspec.reserve((size_t)lsum * n * 8));
  {
    ProfScope ps(ctx, BRA_PROF_OMEGA);
    for (int t = 0; t < T; ++t) {
      const int64_t row0 = nested ? orders[t] - fresh_rows[t] : ctx->spec_off[t];
      int rc = bra_fill_randn(ctx, ctx->omega_spec.as<double>() + row0 * ldt, fresh_rows[t] * ldt, o->seed, (uint64_t)t);
      if (rc) return rc;
    }
  }
  if (!ctx->copy_stream) BRA_CUDA(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
  constexpr int PANELS = 8;
  while ((int)ctx->copy_events.size() < PANELS) {
    cudaEvent_t e;
    BRA_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    ctx->copy_events.push_back(e);
  }
  BRA_CUDA(cudaStreamSynchronize(ctx->stream));        // earlier users of A_stage are done
  const int64_t cper = ((n + PANELS - 1) / PANELS + 255) / 256 * 256;
  int pi = 0;
  for (int64_t c0 = 0; c0 < n; c0 += cper, ++pi) {
    const int64_t cp = (n - c0 < cper) ? n - c0 : cper;
    double* dst = ctx->A_stage.as<double>() + c0 * ld;
    const double* src = A + c0 * lda;
    if (ld == m && lda == m)
      BRA_CUDA(cudaMemcpyAsync(dst, src, (size_t)m * cp * 8, cudaMemcpyHostToDevice, ctx->copy_stream));
    else
      BRA_CUDA(cudaMemcpy2DAsync(dst, (size_t)ld * 8, src, (size_t)lda * 8, (size_t)m * 8, (size_t)cp, cudaMemcpyHostToDevice,
                                 ctx->copy_stream));
    BRA_CUDA(cudaEventRecord(ctx->copy_events[pi], ctx->copy_stream));
    BRA_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->copy_events[pi], 0));
    int rc = bra_gemm_sketch(ctx, ctx->omega_spec.as<double>(), lsum, m, dst, ld, cp, ctx->Bspec.as<double>() + c0 * lsum, lsum);
    if (rc) return rc;
  }
  ctx->spec_rounds = T;
  ctx->spec_ld = lsum;
  ctx->sketch_rows_done = lsum;
  return BRA_OK;
}

int bra_idfact_f64(bra_ctx* ctx, char trans, int64_t m, int64_t n, const double* A, int64_t lda,
                   const bra_opts* opts, const bra_rand* rnd) {
  if (!ctx) return -1;
  int rc = bra_check_fact_args(ctx, trans, m, n, A, lda, opts);
  if (rc) return rc;
  BRA_CUDA(cudaSetDevice(ctx->device));
  const double* dA;
  int64_t dlda;
  rc = bra_stage_A(ctx, trans, m, n, A, lda, opts, rnd, &dA, &dlda);
  if (rc) return rc;
  rc = bra_sketchfact_core(ctx, trans, m, n, dA, dlda, opts, rnd);
  if (rc) return rc;
  BRA_CUDA(cudaStreamSynchronize(ctx->stream));
  return BRA_OK;
}

namespace {
// out[r + c * ldo] = (double)in[r + c * ldi]: one CTA per group of columns, rows coalesced
__global__ void widen_f32_kernel(int64_t m, int64_t n, const float* __restrict__ in, int64_t ldi, double* __restrict__ out,
                                 int64_t ldo) {
  for (int64_t c = blockIdx.y; c < n; c += gridDim.y)
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < ldo; r += (int64_t)gridDim.x * blockDim.x)
      out[r + c * ldo] = (r < m) ? (double)in[r + c * ldi] : 0.0;
}
}  // namespace

// Float32 matrices (the reference's element-type parameter T, src/LowRankApprox.jl:96-118): the kernels of this
// library are FP64 (DMMA), so a Float32 A is widened ONCE on the device -- half the PCIe bytes of a host-side
// conversion -- into a context-owned buffer; the caller hands the returned device pointer to any *_f64 entry point
// and rounds what it fetches.
int bra_widen_f32(bra_ctx* ctx, int64_t m, int64_t n, const float* A, int64_t lda, const double** dA, int64_t* ldd) {
  if (!ctx) return -1;
  BRA_CHECK_ARG(m >= 0, 2, "m");
  BRA_CHECK_ARG(n >= 0, 3, "n");
  BRA_CHECK_ARG(A != nullptr || m == 0 || n == 0, 4, "A");
  BRA_CHECK_ARG(lda >= (m > 1 ? m : 1), 5, "lda");
  BRA_CHECK_ARG(dA != nullptr && ldd != nullptr, 6, "dA / ldd");
  BRA_CUDA(cudaSetDevice(ctx->device));
  const int64_t ld = ((m > 0 ? m : 1) + 1) & ~int64_t(1);
  BRA_CUDA(ctx->A_wide.reserve((size_t)ld * (n > 0 ? n : 1) * 8));
  *dA = ctx->A_wide.as<double>();
  *ldd = ld;
  if (m == 0 || n == 0) return BRA_OK;
  const float* src = A;
  int64_t lds = lda;
  if (!is_device_ptr(A)) {
    BRA_CUDA(ctx->A_stage.reserve((size_t)m * n * 4));
    BRA_CUDA(copy2d(ctx, ctx->A_stage.p, m, A, lda, m, n, 4));
    src = ctx->A_stage.as<float>();
    lds = m;
  }
  dim3 grid((unsigned)std::min<int64_t>((ld + 255) / 256, 64), (unsigned)std::min<int64_t>(n, 148 * 16));
  widen_f32_kernel<<<grid, 256, 0, ctx->stream>>>(m, n, src, lds, ctx->A_wide.as<double>(), ld);
  ctx->launches++;
  BRA_CUDA(cudaGetLastError());
  BRA_CUDA(cudaStreamSynchronize(ctx->stream));
  return BRA_OK;
}

int64_t bra_debug_sketch_rows(bra_ctx* ctx) { return ctx ? ctx->sketch_rows_done : -1; }

// The library's own random numbers, for tests that replay a fast-mode factorization through the oracle:
// bra_debug_randn = `count` values of the Gaussian stream (seed, stream_id) -- row i of a round's fresh Omega rows is
// values [i * ldt, i * ldt + m) with ldt = m rounded up to even; bra_debug_meta = the 8-byte entries bra_fill_meta
// generates (kind 0: +-1.0 signs; 1: distinct subset rows; 2: permutation; 3: SRFT index vector), see sketch_round.
int bra_debug_randn(bra_ctx* ctx, double* host_out, int64_t count, uint64_t seed, uint64_t stream_id) {
  if (!ctx || !host_out || count < 0) return -1;
  BRA_CUDA(cudaSetDevice(ctx->device));
  BRA_CUDA(ctx->scratch2.reserve((size_t)(count > 0 ? count : 1) * 8));
  int rc = bra_fill_randn(ctx, ctx->scratch2.as<double>(), count, seed, stream_id);
  if (rc) return rc;
  BRA_CUDA(cudaMemcpyAsync(host_out, ctx->scratch2.p, (size_t)count * 8, cudaMemcpyDeviceToHost, ctx->stream));
  BRA_CUDA(cudaStreamSynchronize(ctx->stream));
  return BRA_OK;
}
int bra_debug_meta(bra_ctx* ctx, int kind, void* host_out, int64_t count, int64_t range, uint64_t seed, uint64_t stream_id) {
  if (!ctx || !host_out || count < 0 || kind < 0 || kind > 3) return -1;
  BRA_CUDA(cudaSetDevice(ctx->device));
  BRA_CUDA(ctx->scratch2.reserve((size_t)(count > 0 ? count : 1) * 8));
  int rc = bra_fill_meta(ctx, kind, ctx->scratch2.p, count, range, seed, stream_id);
  if (rc) return rc;
  BRA_CUDA(cudaMemcpyAsync(host_out, ctx->scratch2.p, (size_t)count * 8, cudaMemcpyDeviceToHost, ctx->stream));
  BRA_CUDA(cudaStreamSynchronize(ctx->stream));
  return BRA_OK;
}

int bra_get_info(bra_ctx* ctx, bra_info* info) {
  if (!ctx) return -1;
  BRA_CHECK_ARG(info != nullptr, 2, "info");
  std::memset(info, 0, sizeof(*info));
  const FactResult& r = ctx->res;
  info->m = r.m;
  info->n = r.n;
  info->k = r.k;
  info->ksvd = r.ksvd;
  info->rounds = r.rounds;
  for (int t = 0; t < r.rounds && t < BRA_MAX_ROUNDS; ++t) {
    info->orders[t] = r.orders[t];
    info->ks[t] = r.ks[t];
    info->steps[t] = r.steps[t];
  }
  return BRA_OK;
}

int bra_fetch(bra_ctx* ctx, int which, void* dst, int64_t ld) {
  if (!ctx) return -1;
  BRA_CHECK_ARG(dst != nullptr, 3, "dst");
  BRA_CUDA(cudaSetDevice(ctx->device));
  const FactResult& r = ctx->res;
  const int64_t k = r.k, n = r.n, m = r.m;
  switch (which) {
    case BRA_F_P:
      BRA_CUDA(cudaMemcpyAsync(dst, ctx->jpvt.p, (size_t)n * 8, cudaMemcpyDefault, ctx->stream));
      break;
    case BRA_F_T:
      if (!r.have_T) return BRA_ERR_NOTREADY;
      BRA_CHECK_ARG(ld >= (k > 1 ? k : 1), 4, "ld");
      BRA_CUDA(copy2d(ctx, dst, ld, ctx->T.p, r.ldT, k, n - k));
      break;
    case BRA_F_TAU:
      if (r.rounds == 0) return BRA_ERR_NOTREADY;
      BRA_CUDA(cudaMemcpyAsync(dst, ctx->tau.p, (size_t)r.steps[r.rounds - 1] * 8, cudaMemcpyDefault, ctx->stream));
      break;
    case BRA_F_BSKETCH: {
      if (r.rounds == 0) return BRA_ERR_NOTREADY;
      const int64_t l = ctx->cur_rows;          // sketch order (left sketch) or size(op(A), 1) (right sketch)
      BRA_CHECK_ARG(ld >= (l > 1 ? l : 1), 4, "ld");
      if (l <= 0 || n <= 0) break;
      BRA_CUDA(ctx->B2.reserve((size_t)l * n * 8));
      int rc = bra_permute_cols(ctx, ctx->cur_B, ctx->cur_ldb, ctx->B2.as<double>(), l, l, n, ctx->jpvt.as<int64_t>());
      if (rc) return rc;
      BRA_CUDA(copy2d(ctx, dst, ld, ctx->B2.p, l, l, n));
      break;
    }
    case BRA_F_Q:
      if (!r.have_Q) return BRA_ERR_NOTREADY;
      BRA_CHECK_ARG(ld >= (m > 1 ? m : 1), 4, "ld");
      BRA_CUDA(copy2d(ctx, dst, ld, ctx->Q.p, (m + 1) & ~int64_t(1), m, k));
      break;
    case BRA_F_R:
      if (!r.have_R) return BRA_ERR_NOTREADY;
      BRA_CHECK_ARG(ld >= (k > 1 ? k : 1), 4, "ld");
      BRA_CUDA(copy2d(ctx, dst, ld, ctx->Rfull.p, k, k, n));
      break;
    case BRA_F_U:
      if (!r.have_svd || r.svd_vals_only) return BRA_ERR_NOTREADY;
      BRA_CHECK_ARG(ld >= (r.svd_m > 1 ? r.svd_m : 1), 4, "ld");
      BRA_CUDA(copy2d(ctx, dst, ld, ctx->U.p, r.svd_m, r.svd_m, r.ksvd));
      break;
    case BRA_F_S:
      if (!r.have_svd) return BRA_ERR_NOTREADY;
      BRA_CUDA(cudaMemcpyAsync(dst, ctx->S.p, (size_t)r.ksvd * 8, cudaMemcpyDefault, ctx->stream));
      break;
    case BRA_F_VT:
      if (!r.have_svd || r.svd_vals_only) return BRA_ERR_NOTREADY;
      BRA_CHECK_ARG(ld >= (r.ksvd > 1 ? r.ksvd : 1), 4, "ld");
      BRA_CUDA(copy2d(ctx, dst, ld, ctx->Vt.p, r.ksvd, r.ksvd, r.svd_n));
      break;
    default:
      BRA_CHECK_ARG(false, 2, "which");
  }
  BRA_CUDA(cudaStreamSynchronize(ctx->stream));
  return BRA_OK;
}

}  // extern "C"

/* libbrapprox -- C ABI of the B200-native sketch-then-factor path.
 *
 * This header is the drop-in boundary for LowRankApprox.jl's hot path
 * (reference citations are into /root/reference/).  The reference has no
 * plugin interface; its only FFI is the Fortran-style `ccall` into LAPACK
 * (src/lapack.jl:117-139).  The seams replaced here are the Julia-level
 * internal functions
 *     sketch_*(side, trans, A, order, opts) -> B      src/sketch.jl:114,298,530,659
 *     geqp3_adap!(B, opts) -> (p, tau, k)             src/pqr.jl:348
 *     pqrback_postproc(B, p, tau, k, opts)            src/pqr.jl:420
 *     sketchfact(side, trans, A, opts)                src/sketch.jl:52
 *     idfact / pqrfact / psvdfact                     src/id.jl:434, src/pqr.jl:290, src/psvd.jl:238
 *     pheigfact / prange / snorm, snormdiff           src/pheig.jl:276, src/prange.jl:14, src/snorm.jl:14,49
 *
 * Conventions (same as the reference's LAPACK calls): FP64 real, column-major,
 * explicit leading dimensions, int64 dimensions, 1-based index outputs
 * (Julia `Vector{Int}`).  Every pointer argument may be a HOST or a DEVICE
 * pointer (detected with cudaPointerGetAttributes); device-resident inputs are
 * the benchmarked mode.  Every function returns an int status: 0 ok, <0 the
 * number of the invalid argument (LAPACK `info` style), >0 a BRA_ERR_* code;
 * nothing throws or aborts.  There is NO CPU fallback: without a usable B200
 * the calls fail with BRA_ERR_CUDA.
 *
 * A context owns one device, one stream and all workspaces; it is not
 * thread-safe, distinct contexts are independent.
 */
#ifndef BRAPPROX_H
#define BRAPPROX_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BRA_OK 0
#define BRA_ERR_CUDA 1          /* CUDA runtime/driver failure; see bra_last_error */
#define BRA_ERR_UNSUPPORTED 2   /* option combination not built (caller may route to the reference) */
#define BRA_ERR_ROUNDS 3        /* adaptive loop needed more rounds than random inputs supplied */
#define BRA_ERR_INTERNAL 4      /* kernel-side failure (exchange timeout etc.) */
#define BRA_ERR_NOTREADY 5      /* bra_fetch of a factor the last call did not produce */
#define BRA_ERR_COMM 6          /* NCCL failure or libnccl.so.2 not loadable */
#define BRA_ERR_TSLOT 7         /* batched idfact: some block has k > ldT (k_out is valid; enlarge the T slots) */

#define BRA_MAX_ROUNDS 24

/* opts.sketch (src/LowRankApprox.jl:137-138) */
#define BRA_SKETCH_NONE 0
#define BRA_SKETCH_RANDN 1
#define BRA_SKETCH_SPRN 2
#define BRA_SKETCH_SRFT 3
#define BRA_SKETCH_SUB 4

/* opts.retval_mask: the q/r/t substring tests of pqrfact_retval (src/pqr.jl:423-425) */
#define BRA_RET_Q 1
#define BRA_RET_R 2
#define BRA_RET_T 4

/* bra_fetch selectors */
#define BRA_F_P 1        /* int64[n]   pivot permutation p, 1-based (sk = p[0:k], rd = p[k:n]) */
#define BRA_F_T 2        /* double k x (n-k)   interpolation matrix, ld = k */
#define BRA_F_Q 3        /* double m x k       pqrfact Q, ld = m */
#define BRA_F_R 4        /* double k x n       pqrfact R = [R1 | R1 T], ld = k */
#define BRA_F_U 5        /* double m x ksvd    psvdfact U, ld = m */
#define BRA_F_S 6        /* double[ksvd]       psvdfact singular values */
#define BRA_F_VT 7       /* double ksvd x n    psvdfact Vt, ld = ksvd */
#define BRA_F_TAU 8      /* double[k]          Householder scalars of the last QRCP */
#define BRA_F_BSKETCH 9  /* double l x n       last sketch after QRCP, LAPACK layout, ld = l */

typedef struct bra_ctx bra_ctx;

/* POD mirror of LRAOptions' hot-path fields (src/LowRankApprox.jl:77-119).  The
 * three *_samp closures cannot cross a C ABI: the shim evaluates them into the
 * affine pair (samp_a, samp_b), order = samp_a*n + samp_b, or passes explicit
 * per-round orders in bra_rand.orders. */
/* bra_opts.flags.  BRA_OPT_FRESH_SKETCH: with the library's own random numbers, draw every adaptive round's Gaussian
 * Omega independently, like the reference (src/sketch.jl:223-240), instead of nesting the rounds (round t = round t-1's
 * rows + fresh ones, only the new rows are multiplied with A; see bra_debug_sketch_rows).  The environment variable
 * BRA_SKETCH_FRESH=1 sets it for every call. */
#define BRA_OPT_FRESH_SKETCH 1

typedef struct bra_opts {
  double atol;               /* >= 0 */
  double rtol;               /* >= 0; default 5*eps */
  int64_t rank;              /* < 0: unbounded */
  int64_t nb;                /* QRCP block size and first adaptive rank guess; default 32 */
  int32_t sketch;            /* BRA_SKETCH_* */
  int32_t sketch_randn_niter;/* >= 0: Gaussian power iterations (src/sketch.jl:140-149); single-GPU contexts only */
  int32_t sketchfact_adap;   /* default 1 */
  int32_t retval_mask;       /* BRA_RET_* */
  double maxdet_tol;         /* < 0: off; >= 0: strong-RRQR swaps until max|T| <= 1 + tol (src/pqr.jl:444-501) */
  int64_t maxdet_niter;
  int64_t samp_a, samp_b;    /* 0,0 = the reference default for opts.sketch */
  uint64_t seed;             /* fast mode: Philox key */
  int32_t verb;
  int32_t flags;             /* BRA_OPT_* bits; default 0 */
  double pheig_orthtol;      /* >= 0; default sqrt(eps): pheigorth! cluster tolerance (src/pheig.jl:342-364) */
} bra_opts;

/* Per-round random inputs in the order the reference draws them (SURVEY.md
 * Appendix D).  n_rounds == 0 selects the fast mode (device Philox keyed by
 * opts.seed and the round).  Round t uses:
 *   randn: omega[t]  order_t x mA  col-major (ld = order_t)     src/util.jl:4
 *   srft:  d[t] (+-1, length mA), idx[t] (int64, 1-based, length order_t)   src/sketch.jl:339-361
 *   sprn:  perm[t] (int64, 1-based randperm(mA)), s[t] (length mA)          src/sketch.jl:575-579
 *   sub:   r[t] (int64, 1-based, length order_t)                            src/sketch.jl:252
 * where mA is the contracted dimension of op(A). */
typedef struct bra_rand {
  int32_t n_rounds;
  int32_t reserved;
  const double* const* omega;
  const double* const* d;
  const int64_t* const* idx;
  const int64_t* const* perm;
  const double* const* s;
  const int64_t* const* r;
} bra_rand;

typedef struct bra_info {
  int64_t m, n;              /* dims of op(A) */
  int64_t k;                 /* ID rank */
  int64_t ksvd;              /* psvd rank (0 unless psvdfact) */
  int32_t rounds;
  int32_t reserved;
  int64_t orders[BRA_MAX_ROUNDS];
  int64_t ks[BRA_MAX_ROUNDS];
  int64_t steps[BRA_MAX_ROUNDS];   /* pivot steps executed per round (>= ks: blocks run to their end) */
} bra_info;

/* ---- context ----------------------------------------------------------- */
int bra_version(void);
int bra_create(bra_ctx** ctx, int device);
int bra_destroy(bra_ctx* ctx);
const char* bra_last_error(bra_ctx* ctx);
void bra_opts_default(bra_opts* opts);           /* LRAOptions(Float64), src/LowRankApprox.jl:96-119 */
int bra_chkopts(bra_ctx* ctx, const bra_opts* o); /* chkopts!, src/LowRankApprox.jl:133-141 */
uint64_t bra_launch_count(bra_ctx* ctx);         /* kernels launched so far by this ctx */
int bra_sync(bra_ctx* ctx);
void* bra_stream(bra_ctx* ctx);                  /* the ctx's cudaStream_t */

/* ---- stage-wise entry points (parity tests feed each the oracle's input) -- */

/* sketch_randn(:left, trans, A, order) = Omega * op(A)   (src/sketch.jl:129-151)
 * trans = 'n': A is m x n, Omega order x m, B order x n.
 * trans = 'c': B = Omega * A', Omega order x n, B order x m. */
int bra_sketch_randn_f64(bra_ctx* ctx, char trans, int64_t m, int64_t n, const double* A, int64_t lda,
                         int64_t order, const double* Omega, int64_t ldo, double* B, int64_t ldb);

/* The structured sketches, same shapes as bra_sketch_randn_f64, with the random inputs the reference draws
 * (SURVEY.md Appendix D), host or device pointers, indices 1-based int64:
 *   sub  (src/sketch.jl:248-279):  B[i,:] = op(A)[r_i,:],                     r[order]
 *   sprn (src/sketch.jl:571-653):  B[i,:] = sum_t s[off_i+t] op(A)[perm[off_i+t],:],  perm[mA] = randperm, s[mA]
 *   srft (src/sketch.jl:338-522):  sign flip d[mA] (+-1), l x m' reshape, length-l DFT, sampled frequencies
 *                                  idx[order] with (Re, Im) filling row pairs (every other idx entry is unused)
 * mA = size(op(A), 1) is the contracted dimension. */
int bra_sketch_sub_f64(bra_ctx* ctx, char trans, int64_t m, int64_t n, const double* A, int64_t lda, int64_t order,
                       const int64_t* r, double* B, int64_t ldb);
int bra_sketch_sprn_f64(bra_ctx* ctx, char trans, int64_t m, int64_t n, const double* A, int64_t lda, int64_t order,
                        const int64_t* perm, const double* s, double* B, int64_t ldb);
int bra_sketch_srft_f64(bra_ctx* ctx, char trans, int64_t m, int64_t n, const double* A, int64_t lda, int64_t order,
                        const double* d, const int64_t* idx, double* B, int64_t ldb);

/* geqp3_adap!(B, opts) (src/pqr.jl:348-418) on an l x n matrix B, in place:
 * on return B holds R in its upper triangle and the reflectors below (LAPACK
 * layout, columns permuted), jpvt (1-based) the permutation, tau[0:kcap] the
 * scalars, *k the detected rank, *nsteps the pivot steps executed (blocks run
 * to their end, src/pqr.jl:397-414).  kb_trace (optional, host or device,
 * capacity kb_cap) receives the block lengths laqps would have returned. */
int bra_geqp3_adap_f64(bra_ctx* ctx, int64_t l, int64_t n, double* B, int64_t ldb, const bra_opts* opts,
                       int64_t* jpvt, double* tau, int64_t* k, int64_t* nsteps,
                       int32_t* kb_trace, int64_t kb_cap, int64_t* n_blocks);

/* maxdet_t (src/pqr.jl:438-442): T = R[:,0:k]^{-1} R[:,k:n], R is k x n upper trapezoidal. */
int bra_trsolve_T_f64(bra_ctx* ctx, int64_t k, int64_t n, const double* R, int64_t ldr, double* T, int64_t ldt);

/* ---- fused entry points -------------------------------------------------- */

/* idfact(trans, A, opts) (src/id.jl:434-447): results stay in ctx-owned device
 * buffers; read sizes with bra_get_info and copy out with bra_fetch. */
int bra_idfact_f64(bra_ctx* ctx, char trans, int64_t m, int64_t n, const double* A, int64_t lda,
                   const bra_opts* opts, const bra_rand* rnd);

/* pqrfact(trans, A, opts) (src/pqr.jl:290-307), sketched path: idfact, then QR of the skeleton columns
 * (randomised-preconditioned CholeskyQR2) and R = [R1 | R1*T].  Fetch BRA_F_Q (m_op x k), BRA_F_R (k x n_op),
 * BRA_F_P, BRA_F_T.  Q and R equal the reference's Householder factors up to the sign of each column of Q /
 * row of R (Cholesky-QR makes diag(R) > 0; LAPACK's sign is -sign(alpha)). */
int bra_pqrfact_f64(bra_ctx* ctx, char trans, int64_t m, int64_t n, const double* A, int64_t lda,
                    const bra_opts* opts, const bra_rand* rnd);

/* psvdfact(A, opts) (src/psvd.jl:238-272): picks trans = 'n' if m >= n else 'c' like the reference, truncates
 * with psvdrank (src/psvd.jl:301-308).  Fetch BRA_F_U (m x ksvd), BRA_F_S (ksvd), BRA_F_VT (ksvd x n). */
int bra_psvdfact_f64(bra_ctx* ctx, int64_t m, int64_t n, const double* A, int64_t lda, const bra_opts* opts,
                     const bra_rand* rnd);

/* psvdvals(A, opts) (src/psvd.jl:274-290): the singular values of bra_psvdfact_f64 without its vectors (the explicit Q of
 * the skeleton QR, the U and Vt products and the recovery of the rotations are skipped).  Fetch BRA_F_S
 * (bra_get_info().ksvd values); BRA_F_U / BRA_F_VT answer BRA_ERR_NOTREADY. */
int bra_psvdvals_f64(bra_ctx* ctx, int64_t m, int64_t n, const double* A, int64_t lda, const bra_opts* opts,
                     const bra_rand* rnd);

/* Host destinations for the factors of the NEXT bra_psvdfact_f64 on this context (one shot; cleared when that call
 * returns).  U: column-major, ldu >= m, room for ucols columns; S: room for scap values; Vt: leading dimension ldvt
 * (usable when the rank found is <= ldvt).  A factor that fits is copied to the host inside the call, on the stream
 * that produced it -- the copy of the factor that is ready first overlaps the rest of the computation -- and
 * bra_psvd_outputs_done() reports which ones were written (bit 0: U, bit 1: S, bit 2: Vt); the others are fetched with
 * bra_fetch as usual.  Any pointer may be NULL.  Only page-locked (pinned) host memory is written inside the call (a copy
 * into pageable memory would block the calling thread in the middle of the factorization); pageable destinations are
 * ignored here and filled by bra_fetch. */
int bra_psvd_set_outputs(bra_ctx* ctx, double* U, int64_t ldu, int64_t ucols, double* S, int64_t scap, double* Vt,
                         int64_t ldvt);
int bra_psvd_outputs_done(bra_ctx* ctx);

/* pheigfact(A, opts) (src/pheig.jl:276-296) for a real symmetric n x n A (-3: "matrix must be Hermitian", :279):
 * idfact, QR of [I; T'] (one Cholesky pass), the k x k eigenproblem of R (A[sk,sk] R') on the device (one-sided
 * Jacobi + Rayleigh quotients), truncation with pheigrank (:322-341).  Fetch BRA_F_S (kk eigenvalues, ascending,
 * negative part first) and BRA_F_U (n x kk eigenvectors); bra_get_info().ksvd = kk.  pheigvals = fetch BRA_F_S only. */
int bra_pheigfact_f64(bra_ctx* ctx, int64_t n, const double* A, int64_t lda, const bra_opts* opts, const bra_rand* rnd);

/* sketchfact(side, trans, A, opts) (src/sketch.jl:52-66; drivers :213-240, 313-330, 545-562, 674-690): the
 * early-terminating pivoted QR of the SKETCH of op(A).  side 'l': B = S op(A) (order x n_op) -- the first stage of
 * idfact / pqrfact / psvdfact; fetch BRA_F_P, BRA_F_T, BRA_F_TAU and BRA_F_BSKETCH (LAPACK layout: R = triu of the first
 * k rows, reflectors below the diagonal).  side 'r': B = op(A) S (m_op x order), the range-finder form behind prange;
 * fetch BRA_F_Q (m_op x k, the Householder Q up to the sign of each column), BRA_F_P (order entries), BRA_F_TAU,
 * BRA_F_BSKETCH (m_op x order), and BRA_F_T when maxdet_tol >= 0.  bra_get_info: k, rounds, orders.  Errors: -2 side
 * (sketchfact_chkargs, :80-84), the bra_idfact_f64 argument numbers shifted by one, -8 for sketch = :none. */
int bra_sketchfact_f64(bra_ctx* ctx, char side, char trans, int64_t m, int64_t n, const double* A, int64_t lda,
                       const bra_opts* opts, const bra_rand* rnd);

/* CUR(A, rows, cols) / HermCUR(A, cols) (src/cur.jl:85-109): the factors of a CUR decomposition from its index sets
 * (1-based, k entries each; what curfact returns).  C = A[:, cols] (BRA_F_Q, m x k), R = A[rows, :] (BRA_F_R, k x n), and
 * the k x k core on the device (one-sided Jacobi):
 *   hermitian == 0:  U, s, V = svd!(C[rows, :]);  U2 = PartialSVD(V, 1 ./ s, U'):
 *                    BRA_F_U = V (k x k), BRA_F_S = 1 ./ s (s descending), BRA_F_VT = U' (k x k);
 *   hermitian != 0:  F = eigen!(Hermitian(C[cols, :])) (rows is ignored, A must be square):
 *                    BRA_F_S = 1 ./ F.values (values ascending), BRA_F_U = F.vectors (k x k); no R. */
int bra_cur_f64(bra_ctx* ctx, int64_t m, int64_t n, const double* A, int64_t lda, int64_t k, const int64_t* rows1,
                const int64_t* cols1, int hermitian);

/* Rows of Omega that the last Gaussian-sketch factorization multiplied with op(A) (diagnostic).  With the library's own
 * Omega the rounds of the adaptive loop are NESTED -- round t's Omega is round t-1's followed by fresh rows, so only the
 * new rows are multiplied and the total is max(order) instead of the reference schedule's sum(order)
 * (src/sketch.jl:223-240 draws every round afresh); caller-supplied Omegas, power iterations and BRA_SKETCH_FRESH=1
 * keep the round-by-round products. */
int64_t bra_debug_sketch_rows(bra_ctx* ctx);
/* The library's own random numbers (test hooks: a fast-mode factorization can be replayed through a CPU oracle on the
 * very same inputs).  bra_debug_randn: `count` values of the Gaussian stream (seed, stream_id); row i of the fresh Omega
 * rows of adaptive round t is values [i*ldt, i*ldt + m) of stream t, ldt = m rounded up to even.  bra_debug_meta: the
 * 8-byte entries of kind 0 (+-1.0 SRFT signs), 1 (distinct subset rows, int64), 2 (permutation, int64), 3 (SRFT index
 * vector, int64) for (seed, stream_id = round). */
int bra_debug_randn(bra_ctx* ctx, double* host_out, int64_t count, uint64_t seed, uint64_t stream_id);
int bra_debug_meta(bra_ctx* ctx, int kind, void* host_out, int64_t count, int64_t range, uint64_t seed, uint64_t stream_id);

/* Float32 matrices.  The reference is generic in the element type T (LRAOptions(T), src/LowRankApprox.jl:96-118;
 * pqrfact / idfact / psvdfact / ... take AbstractMatOrLinOp{T}); the kernels here are FP64.  bra_widen_f32 uploads a
 * Float32 A (host or device, column-major, lda in elements) and widens it on the device into a context-owned FP64 copy;
 * *dA / *ldd are then passed as the device-resident `A` / `lda` of any *_f64 entry point, and the caller rounds the
 * fetched factors to Float32.  The copy stays valid until the next bra_widen_f32 on the same context.  The
 * factorization itself runs in FP64, so with T = Float32 defaults (rtol = 5 eps(Float32)) the results agree with the
 * reference's Float32 path to Float32 accuracy, not bit for bit.  ComplexF32 / ComplexF64 are not built. */
int bra_widen_f32(bra_ctx* ctx, int64_t m, int64_t n, const float* A, int64_t lda, const double** dA, int64_t* ldd);

/* prange(trans, A, opts) (src/prange.jl:14-62): an orthonormal basis of the range of A (trans 'n'), of A' ('c') or of
 * both ('b', square A; a Hermitian A falls back to 'n', :26).  sketch = :none -> pqrfact(op(A))[:Q] (:50-52); :sub ->
 * prange_sub (:64-77); otherwise the pivoted QR of the right-hand sketch B = op(A) S, sketchfact(:right, trans, A, opts)
 * (src/sketch.jl:52-66), S on the same random inputs as the left-hand sketch of op(A)' (the contracted dimension is
 * size(op(A), 2)).  rnd2: the inputs of the second sketch of trans 'b' (the reference factors A' first, then A).
 * Result: Q = BRA_F_Q, info.m x info.k, equal to the reference's Householder Q up to the sign of each column. */
int bra_prange_f64(bra_ctx* ctx, char trans, int64_t m, int64_t n, const double* A, int64_t lda, const bra_opts* opts,
                   const bra_rand* rnd, const bra_rand* rnd2);

/* snorm(A - L R) (snorm / snormdiff, src/snorm.jl:14-53): randomised power iteration on DEVICE-resident operands,
 * A m x n, L m x k, R k x n (k = 0: snorm(A); a Hermitian A then takes one product per iteration, :33-35).  Stops when
 * |s - s_prev| <= max(opts->atol, s_prev * opts->rtol) or after niter_max iterations (LRAOptions.snorm_niter).
 * x0: optional device start vector (the reference's crandn(n)); NULL = device Philox keyed by opts->seed. */
int bra_snorm_f64(bra_ctx* ctx, int64_t m, int64_t n, const double* A, int64_t lda, int64_t k, const double* L, int64_t ldl,
                  const double* R, int64_t ldr, const bra_opts* opts, int64_t niter_max, const double* x0, double* result,
                  int64_t* niter_out);

/* ---- batched idfact of independent blocks (BASELINE config 5; additive: the reference loops idfact) --------
 * nblocks column-major m x n blocks, block b at A + b*strideA (leading dimension lda), all DEVICE resident.
 * opts as for idfact with sketch = :sprn.  Random inputs in reference order per block (src/sketch.jl:575-579):
 * perm (1-based randperm(m)) and s (m weights), block b at perm + b*perm_stride / s + b*s_stride (stride 0 shares
 * one draw); perm = s = NULL selects the fast mode (one library-drawn (perm, s) per call, shared by the batch).
 * Outputs (device): k_out[b]; p_out[b*n .. b*n+n) 1-based pivots (sk = first k_b, rd = rest); T_b (k_b x (n-k_b))
 * at T_out + b*strideT with leading dimension ldT.  One fused kernel (sketch -> QRCP -> T, one CTA per block) runs
 * the first adaptive round; blocks with k_b >= nb continue through the general path (bra_batched_unfinished tells
 * how many).  Returns BRA_ERR_TSLOT if some k_b > ldT (k_out and p_out are valid, that block's T is not stored). */
int bra_idfact_batched_f64(bra_ctx* ctx, int64_t nblocks, int64_t m, int64_t n, const double* A, int64_t lda,
                           int64_t strideA, const bra_opts* opts, const int64_t* perm, int64_t perm_stride,
                           const double* s, int64_t s_stride, int64_t* k_out, int64_t* p_out, double* T_out,
                           int64_t ldT, int64_t strideT);
int64_t bra_batched_unfinished(bra_ctx* ctx);
/* Cycles CTA 0 spent in the last batched launch (diagnostic), 8 entries: random tables, waiting for tiles, QRCP,
 * outputs + T, tile issue, sketch gather, hand-over to registers, unused. */
int bra_debug_batched_phases(bra_ctx* ctx, int64_t* out8);

/* ---- multi-GPU: row-sharded tall matrices (BASELINE config 4) ----------------------
 * One process per GPU, one ctx per process.  Rank 0 calls bra_comm_unique_id and ships the 128 bytes to the other
 * ranks by any means (the Python host uses torch.distributed); every rank then calls bra_comm_init, which builds
 * an NCCL communicator bound to the ctx.  After bra_set_row_shard(row0, m_global) the fused entry points
 * (bra_idfact_f64 / bra_pqrfact_f64 / bra_psvdfact_f64 with trans = 'n', sketch = randn) take the LOCAL row block
 * (m_local x n): the sketch is all-reduced once per adaptive round, the CholeskyQR2 Gram matrices once per pass;
 * p, k, T, R, S, Vt come out replicated, Q / U hold the local rows.  Fast-mode Omega is keyed by the GLOBAL row
 * index, so the factorization does not depend on the number of ranks beyond summation order. */
int bra_comm_unique_id(void* id128);
int bra_comm_init(bra_ctx* ctx, const void* id128, int rank, int world);
int bra_comm_destroy(bra_ctx* ctx);
int bra_set_row_shard(bra_ctx* ctx, int64_t row0, int64_t m_global);
uint64_t bra_collective_count(bra_ctx* ctx);

int bra_get_info(bra_ctx* ctx, bra_info* info);
int bra_fetch(bra_ctx* ctx, int which, void* dst, int64_t ld);

/* ---- per-stage device timing (CUDA events recorded on the ctx stream around each stage) ---- */
#define BRA_PROF_OMEGA 0     /* Omega generation / transpose */
#define BRA_PROF_GEMM 1      /* sketch GEMM kernel (TMA + DMMA) */
#define BRA_PROF_SPLITK 2    /* deterministic split-K reduction */
#define BRA_PROF_QRCP 3      /* persistent QRCP kernel */
#define BRA_PROF_GATHER 4    /* R gather / permutations */
#define BRA_PROF_TRSOLVE 5   /* T = R11^-1 R12 */
#define BRA_PROF_TAIL 6      /* pqr / psvd tail kernels */
#define BRA_PROF_SKETCH_OTHER 7 /* srft / sprn / sub sketch kernels */
#define BRA_PROF_SVD 8       /* psvd core: k x k Jacobi SVD */
#define BRA_PROF_QR 9        /* skeleton / Z CholeskyQR2 (Gram GEMMs, Cholesky, triangular solves) */
#define BRA_PROF_TAILGEMM 10 /* GEMMs of the pqr / psvd tails (same TMA + DMMA kernel) */
#define BRA_PROF_COMM 11     /* NCCL all-reduces (row-sharded sketch, Gram matrices) */
#define BRA_PROF_BATCHED 12  /* fused one-CTA-per-block batched idfact kernel */
#define BRA_PROF_NTAGS 13
int bra_profile_enable(bra_ctx* ctx, int on);            /* also clears the accumulators */
/* accumulated milliseconds and span counts per tag since the last enable; syncs the stream */
int bra_profile_read(bra_ctx* ctx, double* ms, int64_t* calls);

/* ---- diagnostics ---------------------------------------------------------- */
/* FP64 peak probes used as roofline denominators (bench.py): a register-resident
 * DMMA loop and a DFMA loop; returns TFLOP/s in out[0..4]:
 * m8n8k4, DFMA, m16n8k4, m16n8k8, m16n8k16 (all lower to DMMA.8x8x4 on sm_100a). */
int bra_probe_fp64_peak(bra_ctx* ctx, double* out);
/* Average latency (microseconds) of one LL all-gather exchange across `ctas` CTAs. */
int bra_probe_exchange_latency(bra_ctx* ctx, int ctas, int iters, double* usec);

/* Exchange design probe: mode 0 = push `hw` words to every inbox, mode 1 = two hops through `leaders`. */
int bra_probe_exchange2(bra_ctx* ctx, int ctas, int hw, int mode, int leaders, int iters, double* usec);

/* Kilo-cycles CTA 0 spent per phase of the last QRCP launch: candidate merge + header push, dlarfg on its own
 * candidate + record push, header gather, winner-vector fetch, rank-1 update + norm downdate, (unused)
 * (clock64 deltas; diagnostic only). */
int bra_debug_qrcp_phases(bra_ctx* ctx, int32_t* out6);
/* Sweeps the last psvd core (k x k Jacobi SVD) needed. */
int bra_debug_jacobi_sweeps(bra_ctx* ctx);
/* test hook: pheigorth! (src/pheig.jl:342-364) on host arrays, in place (vals ascending, V rows x kk, ld) */
int bra_debug_pheigorth(bra_ctx* ctx, const double* vals, double* V, int64_t ld, int rows, int kk, double orthtol);
/* Skeleton QRs (the QR of A[:, sk] in pqrfact / psvdfact / prange, src/pqr.jl:297-305) this context had to redo with a
 * fresh Gaussian preconditioner after the Gram Cholesky of the first attempt broke down (cumulative). */
int bra_debug_skeleton_retries(bra_ctx* ctx);
/* column swaps made by the last maxdet post-processing (maxdet_swapcols!, src/pqr.jl:444-478) */
int64_t bra_debug_maxdet_swaps(bra_ctx* ctx);
/* Kilo-cycles thread 0 spent in the last Jacobi launch: panel load, round sync, panel store, grid barrier,
 * dot products, shuffles, rotation scalars, rotation apply (out8[0..7]). */
int bra_debug_jacobi_phases(bra_ctx* ctx, int32_t* out8);
/* Same, for every CTA of the last QRCP launch: out[cta*8 + phase], ctas <= 160. */
int bra_debug_qrcp_phases_all(bra_ctx* ctx, int32_t* out, int ctas);
/* diagnostic: clock64 stamps [ctas][16][16] of the pivot step named by BRA_QRCP_TS_STEP in the last QRCP launch */
int bra_debug_qrcp_trace(bra_ctx* ctx, int64_t* out, int ctas);

#ifdef __cplusplus
}
#endif
#endif /* BRAPPROX_H */

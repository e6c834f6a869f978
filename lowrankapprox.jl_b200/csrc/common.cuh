// Shared declarations for libbrapprox (B200 / sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda.h>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>
#include "../../include/brapprox.h"

#define BRA_CUDA(expr)                                                            \
  do {                                                                            \
    cudaError_t _e = (expr);                                                      \
    if (_e != cudaSuccess) {                                                      \
      ctx->set_error(std::string(#expr) + ": " + cudaGetErrorString(_e));         \
      return BRA_ERR_CUDA;                                                        \
    }                                                                             \
  } while (0)

#define BRA_CHECK_ARG(cond, argno, msg)                                           \
  do {                                                                            \
    if (!(cond)) {                                                                \
      ctx->set_error(std::string("invalid argument ") + #argno + ": " + msg);     \
      return -(argno);                                                            \
    }                                                                             \
  } while (0)

// A growable device buffer owned by the context (never freed behind the caller's back).
struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
  cudaError_t reserve(size_t bytes) {
    if (bytes <= cap) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    size_t want = bytes + bytes / 8 + 256;
    cudaError_t e = cudaMalloc(&p, want);
    if (e == cudaSuccess) cap = want;
    return e;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
  }
  template <class T> T* as() { return reinterpret_cast<T*>(p); }
};

// Result of the last fused factorization, resident on the device (two-phase protocol).
struct FactResult {
  int64_t m = 0, n = 0;        // dims of op(A)
  int64_t k = 0;               // ID rank
  int64_t ksvd = 0;            // psvd rank after truncation
  int rounds = 0;
  int64_t orders[BRA_MAX_ROUNDS];
  int64_t ks[BRA_MAX_ROUNDS];
  int64_t steps[BRA_MAX_ROUNDS];
  int64_t ldT = 0;             // leading dimension of ctx->T (k rounded up to even: keeps the TMA GEMM path open)
  int64_t svd_m = 0, svd_n = 0;  // dims of the ORIGINAL A for psvdfact's U (svd_m x ksvd) and Vt (ksvd x svd_n)
  bool have_T = false, have_Q = false, have_R = false, have_svd = false;
  bool svd_vals_only = false;  // psvdvals: only BRA_F_S is valid
  bool maxdet_done = false;    // maxdet swapped columns: R11 no longer belongs to the skeleton (tails must not use it)
};

constexpr size_t BRA_HPIN_BYTES = 256 * 1024;

struct bra_ctx {
  int device = 0;
  int num_sms = 0;
  int smem_optin = 0;
  cudaStream_t stream = nullptr;
  std::string err;
  uint64_t launches = 0;       // kernels launched by this ctx (bench's gpu_launches)

  // workspaces
  DevBuf A_stage;              // staging for host-resident A
  DevBuf A_wide;               // a Float32 A widened to FP64 (bra_widen_f32)
  DevBuf omega_t;              // Omega^T, K-major  [order][m]
  DevBuf omega_in;             // staging for a host-resident Omega
  DevBuf B;                    // sketch, l x n (col-major, ld = l)
  DevBuf B2;                   // permuted copy / scratch
  DevBuf Bnew, Braw;           // nested Gaussian sketches: the new rows of a round, the raw (unfactored) rows so far
  // one-shot host destinations of the next psvdfact (bra_psvd_set_outputs): each factor is copied out on the stream
  // that produced it, as soon as it exists, instead of by bra_fetch after the call
  double* out_U = nullptr; double* out_S = nullptr; double* out_Vt = nullptr;
  int64_t out_ldu = 0, out_ucols = 0, out_ldvt = 0, out_scap = 0;
  int out_done = 0;               // bit 0: U, 1: S, 2: Vt written by the last psvdfact
  int64_t sketch_rows_done = 0;   // rows of Omega multiplied with op(A) by the last factorization (bra_debug_sketch_rows)
  DevBuf partial;              // split-K partial sums
  DevBuf vn1, vn2, lpos, fpend; // QRCP per-column state
  DevBuf rec;                  // LL exchange records
  DevBuf jpvt, tau, rdiag, info, kbtrace;
  DevBuf R11, T;               // k x k, k x (n-k)
  DevBuf C, Q, R1, Rfull;      // pqr tail
  DevBuf W, G, U, S, Vt, Z;    // psvd tail
  DevBuf scratch, scratch2, scratch3;
  DevBuf aux_in1, aux_in2;     // staged random inputs (d, idx, perm, s, r)
  DevBuf jwork;                // Jacobi SVD: grid barrier, per-sweep flags
  DevBuf rinv, yt;             // CholeskyQR: explicit triangular inverse, transposed panels
  DevBuf tritmp;               // blocked triangular inverse: B C^{-1} scratch
  DevBuf cholscr;              // blocked Cholesky: inverse of the current diagonal block
  DevBuf Bq;                   // power iteration: the sketch on the other side of A
  // side lane: an independent chain of small kernels (psvdfact: Z'Z -> R_z -> R_z^{-1}) runs on a second stream next to
  // the main chain; the helpers that own a workspace or a status word pick the lane's copy (bra_lane_* in tail.cu)
  int lane = 0;
  cudaStream_t side_stream = nullptr, lane_saved = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  DevBuf partial_l1, cholscr_l1, tritmp_l1, G_l1;
  DevBuf Qt_l1, Out_l1;         // psvdfact: Q' and U' formed on the side lane
  DevBuf& ws_partial() { return lane ? partial_l1 : partial; }
  DevBuf& ws_cholscr() { return lane ? cholscr_l1 : cholscr; }
  DevBuf& ws_tritmp() { return lane ? tritmp_l1 : tritmp; }
  DevBuf Bt, Bcat;             // prange: the tall right-hand sketch B = op(A) S; [B_r[:,p_r] B_c[:,p_c]] of the two-sided form
  // host-resident A: the upload is pipelined with the sketch products of the first adaptive rounds (api.cu: bra_stage_A)
  cudaStream_t copy_stream = nullptr;
  std::vector<cudaEvent_t> copy_events;
  DevBuf omega_spec, Bspec;    // stacked Omega^T / sketches of the speculative rounds
  int spec_rounds = 0;         // rounds whose sketch is already in Bspec (0: none)
  int64_t spec_ld = 0, spec_off[BRA_MAX_ROUNDS] = {0};
  double* cur_B = nullptr;     // sketch of the current round and its leading dimension
  int64_t cur_ldb = 0;
  int64_t cur_rows = 0;        // rows of cur_B (the sketch order for a left sketch, size(op(A), 1) for a right one)
  int A_sym_state = 0;         // power iteration: 0 unknown, 1 A == A' exactly, -1 not (checked once per factorization)
  int last_jacobi_sweeps = 0;
  int jacobi_kcycles[8] = {0};
  std::vector<unsigned char> h_meta;  // host scratch for fast-mode index/sign generation
  DevBuf At;                   // transposed copy of A for the (:left,:c) SRFT
  bool At_valid = false;
  DevBuf Apanels;              // row-panel copy of a tall A for the Gaussian sketch (made once per factorization)
  int Apanels_state = 0;       // 0: not tried, 1: valid, -1: not used (small A or not enough memory)
  // multi-GPU (comm.cu): NCCL communicator over the ranks that hold the row blocks of one tall matrix
  void* nccl_comm = nullptr;
  int rank = 0, world = 1;
  int64_t shard_row0 = 0, shard_m_global = 0;   // this rank's first global row / total rows (0: not sharded)
  uint64_t collectives = 0;
  int64_t batched_unfinished = 0;   // blocks of the last batched call that needed rounds beyond the fused one
  int64_t last_maxdet_swaps = 0;    // column swaps done by the last maxdet post-processing
  int start_round = 0;              // adaptive loop resumes at this round (batched fallback)
  int skeleton_retries = 0;         // skeleton QRs redone with a fresh preconditioner after a Cholesky breakdown
  int gemm_tag = BRA_PROF_GEMM; // profiling tag the GEMM launchers record under (tails switch it)
  uint32_t rec_epoch = 1;
  size_t rec_zeroed = 0;
  int32_t* h_info = nullptr;   // pinned, 16 ints
  unsigned char* h_pin = nullptr;   // pinned scratch for small read-backs / uploads (BRA_HPIN_BYTES)

  FactResult res;

  // optional per-stage device timing (CUDA events on the launching stream); see bra_profile_*
  bool prof_on = false;
  struct ProfSpan { int tag; cudaEvent_t e0, e1; };
  std::vector<ProfSpan> prof_spans;
  std::vector<cudaEvent_t> prof_pool;
  double prof_ms[BRA_PROF_NTAGS] = {0};
  int64_t prof_calls[BRA_PROF_NTAGS] = {0};
  cudaEvent_t prof_event() {
    if (!prof_pool.empty()) { cudaEvent_t e = prof_pool.back(); prof_pool.pop_back(); return e; }
    cudaEvent_t e; cudaEventCreate(&e); return e;
  }
  void prof_begin(int tag) {
    if (!prof_on || lane) return;      // side-lane work hides behind the main chain: not a stage of its own
    ProfSpan sp{tag, prof_event(), prof_event()};
    cudaEventRecord(sp.e0, stream);
    prof_spans.push_back(sp);
  }
  void prof_end() {
    if (!prof_on || lane || prof_spans.empty()) return;
    // close the most recent open span
    cudaEventRecord(prof_spans.back().e1, stream);
  }

  void set_error(const std::string& s) { err = s; }
};

struct ProfScope {
  bra_ctx* c;
  ProfScope(bra_ctx* ctx, int tag) : c(ctx) { c->prof_begin(tag); }
  ~ProfScope() { c->prof_end(); }
};

static inline bool is_device_ptr(const void* p) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

// ---- kernels' host launchers (defined in the .cu files) -------------------

// qrcp.cu
struct QrcpOut {
  int k, nsteps, nblocks, status;
};
int bra_qrcp_run(bra_ctx* ctx, double* B, int64_t ldb, int l, int64_t n, int kcap, int nb,
                 double atol, double rtol, QrcpOut* out, bool nopivot = false);
// qrcp_blocked.cu: blocked dlaqps for tall matrices (trailing update on DMMA)
bool bra_qrcp_blocked_ok(int64_t l, int64_t n, int nb, int num_sms);
int bra_qrcp_blocked_run(bra_ctx* ctx, double* B, int64_t ldb, int64_t l, int64_t n, int kcap, int nb, double atol,
                         double rtol, QrcpOut* out);
int bra_permute_cols(bra_ctx* ctx, const double* src, int64_t lds, double* dst, int64_t ldd,
                     int64_t rows, int64_t n, const int64_t* jpvt1);
int bra_gather_R(bra_ctx* ctx, const double* B, int64_t ldb, int64_t n, int k,
                 const int64_t* jpvt1, double* R11, double* R12, int64_t ld12);

// sketch_randn.cu
int bra_transpose_omega(bra_ctx* ctx, const double* Om, int64_t ldo, int64_t l, int64_t m, double* Omt);
int bra_fill_randn(bra_ctx* ctx, double* dst, int64_t count, uint64_t seed, uint64_t stream_id);
int bra_fill_randn_rows(bra_ctx* ctx, double* dst, int64_t ldt, int64_t order, int64_t row0, int64_t ldg, uint64_t seed,
                        uint64_t stream_id);
int bra_gemm_sketch(bra_ctx* ctx, const double* Omt, int64_t l, int64_t m, const double* A, int64_t lda,
                    int64_t n, double* B, int64_t ldb);
int bra_gemm_tn(bra_ctx* ctx, const double* X, int64_t ldx, int64_t l, int64_t m, const double* Y, int64_t ldy,
                int64_t n, double* C, int64_t ldc);
bool bra_gemm_tma_ok(const double* A, int64_t lda, int64_t m, int64_t n);
bool bra_gemm_wants_panels(const double* A, int64_t lda, int64_t m, int64_t n);
int64_t bra_panel_bytes(int64_t m, int64_t n);
int bra_repack_panels(bra_ctx* ctx, const double* A, int64_t lda, int64_t m, int64_t n, double* P);
int bra_gemm_sketch_panels(bra_ctx* ctx, const double* Omt, int64_t l, int64_t m, const double* P, int64_t n, double* B,
                           int64_t ldb);
bool bra_make_map_3d_f64(CUtensorMap* map, const double* base, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t s1,
                         uint64_t s2, uint32_t b0, uint32_t b1);

// comm.cu
int bra_allreduce_sum_f64(bra_ctx* ctx, double* buf, int64_t count);

// sketch_other.cu
int bra_sketch_sub(bra_ctx* ctx, char trans, const double* A, int64_t lda, int64_t mA, int64_t nA, int64_t order,
                   const int64_t* r1_dev, double* B, int64_t ldb);
int bra_sketch_sprn(bra_ctx* ctx, char trans, const double* A, int64_t lda, int64_t mA, int64_t nA, int64_t order,
                    const int64_t* perm1_dev, const double* s_dev, double* B, int64_t ldb);
int bra_sketch_srft(bra_ctx* ctx, const double* A, int64_t lda, int64_t mA, int64_t nA, int64_t order,
                    const double* d_dev, const int64_t* idx1_dev, double* B, int64_t ldb);
int bra_fill_meta(bra_ctx* ctx, int kind, void* dst_dev, int64_t count, int64_t range, uint64_t seed, uint64_t stream_id);
int bra_splitk_reduce(bra_ctx* ctx, const double* part, int64_t split_stride, int splits, int64_t l, int64_t n,
                      double* out, int64_t ldo);

// tail.cu (pqrfact / psvdfact tails)
int bra_gather_cols(bra_ctx* ctx, char trans, const double* A, int64_t lda, int64_t mC, int64_t k,
                    const int64_t* idx1, double* C, int64_t ldc);
int bra_transpose(bra_ctx* ctx, const double* src, int64_t lds, int64_t rows, int64_t cols, double* dst, int64_t ldd);
int bra_trsolve_right_upper(bra_ctx* ctx, int64_t rows, int k, const double* R, int64_t ldr, double* Y, int64_t ldy);
int bra_cholesky_upper(bra_ctx* ctx, int k, double* G, int64_t ldg, double* Rout, int64_t ldr);
int bra_cholqr2(bra_ctx* ctx, int64_t rows, int k, double* Y, int64_t ldy, const double* Rpre, double* Rout,
                bool rows_sharded = false, bool defer_last_solve = false);
int bra_jacobi_svd(bra_ctx* ctx, int k, double* X, int64_t ldx, double* J, int64_t ldj, double* sigma_host,
                   int* order_host, bool* skip_J = nullptr);
extern "C" {
int bra_sketchfact_core(bra_ctx* ctx, char trans, int64_t m, int64_t n, const double* dA, int64_t lda,
                        const bra_opts* o, const bra_rand* rnd);
int bra_check_fact_args(bra_ctx* ctx, char trans, int64_t m, int64_t n, const double* A, int64_t lda,
                        const bra_opts* opts);
int bra_prange_core(bra_ctx* ctx, char trans, int64_t m, int64_t n, const double* dA, int64_t lda, const bra_opts* o,
                    const bra_rand* rnd, bool with_maxdet);
int bra_is_symmetric_dev(bra_ctx* ctx, int64_t n, const double* dA, int64_t lda, int* sym);
int bra_stage_A(bra_ctx* ctx, char trans, int64_t m, int64_t n, const double* A, int64_t lda, const bra_opts* opts,
                const bra_rand* rnd, const double** dA, int64_t* dlda);
}
int bra_gemm_generic(bra_ctx* ctx, const double* Om, int64_t osi, int64_t osk, const double* A, int64_t sk,
                     int64_t sj, int64_t l, int64_t n, int64_t K, double* C, int64_t ldc);

// maxdet.cu
int bra_maxdet_swapcols(bra_ctx* ctx, int k, int64_t ncols, double* T, int64_t ld, int64_t* jpvt, double tol,
                        int64_t niter_max, int64_t* nswaps);

// tail.cu helpers shared with the sketch driver
int bra_set_identity(bra_ctx* ctx, int k, double* J, int64_t ldj);
int bra_cholesky_upper(bra_ctx* ctx, int k, double* G, int64_t ldg, double* Rout, int64_t ldr);
int bra_chol_status_reset(bra_ctx* ctx);
int bra_chol_status(bra_ctx* ctx);

// trsolve.cu
int bra_trsolve_upper(bra_ctx* ctx, int k, int64_t nrhs, const double* R11, int64_t ldr, double* X, int64_t ldx);
int bra_tri_inverse_upper(bra_ctx* ctx, int k, const double* R, int64_t ldr, double* Rinv, int64_t ldx);

// Early-terminating Householder QRCP of a TALL matrix, BLOCKED: LAPACK's dlaqps exactly as the reference drives it
// (geqp3_adap_main!, src/pqr.jl:361-418; _LAPACK.laqps!, src/lapack.jl:117-139).  This is the north star's "blocked
// Householder trailing update on the tensor cores":
//   * per pivot step ONE read-only pass over the trailing matrix -- the new column of F, F[:, k] = tau A[rk:, :]' v (plus
//     the V'v correction) -- and the pivot ROW only is brought up to date; the reflectors of up to nb steps accumulate;
//   * at the block end the trailing update A22 -= V F' runs on the FP64 tensor cores (mma.sync.m8n8k4.f64 = DMMA),
//     one read and one write of the trailing matrix per BLOCK instead of per step.
// Used for the pivoted QRs whose rows do not fit on chip: sketch = :none (pqrfact_none, src/pqr.jl:323-327) and the tall
// right-hand sketches of prange / sketchfact(:right) (src/prange.jl:14-62).  The short sketches of the headline path
// (l <= 576 rows) stay on qrcp_fast.cu: there the slab lives in tensor / shared memory, the rank-1 update hides behind the
// pivot exchange, and deferring it would only add the V'v correction to the latency chain of every step.
//
// Semantics (verified against the real dlaqps through the oracle): first-maximum pivot, the norms travel with their
// columns, LAWN-176 downdate with tol3z = sqrt(2^-53), a flagged column ends the block after the current step, flagged
// norms are recomputed from the UPDATED column after the block's trailing update, rank test at block ends only.
// Structure: one persistent cooperative kernel; CTA c owns n/G physical columns (never swapped: each carries its logical
// LAPACK position); its rows of F live in shared memory for the whole block; two grid barriers per step.
#include "common.cuh"
#include <cstdlib>
#include "qrcp_common.cuh"

namespace {

constexpr int QB_THREADS = 512;
constexpr int QB_WARPS = QB_THREADS / 32;
constexpr int QB_NB = 32;             // largest block size (reflectors accumulated): opts.nb <= 32
constexpr int QB_FLD = 36;            // row stride of F in shared memory (= 4 mod 16: conflict-free fragment loads)
constexpr int QB_CPW = 8;             // columns per warp  -> at most 128 columns per CTA
constexpr int QB_CPC = QB_CPW * QB_WARPS;
constexpr int QB_MAXG = 160;
constexpr int QB_VCH = 8192;          // rows of the Householder vector staged in shared memory at a time
constexpr unsigned QB_SPIN = 1u << 22;      // ~2 s of polling: a lost CTA makes every other one time out and report

struct __align__(16) QbCand {
  double key;     // candidate norm (< 0: none)
  int lp;         // logical position
  int phys;       // physical column
  int flag;       // a column of this CTA was flagged in the step just done
  int pad[3];
};

struct QbParams {
  double* A;
  int64_t lda, m, n;
  int kcap, nb;
  double atol, rtol;
  int cpc;
  QbCand* cand;        // [2][G]
  double* hdr;         // [0] tau, [1] beta, [8 .. 8+nb) aux
  unsigned* bar;       // grid barrier counter (zeroed before launch)
  int64_t* jpvt;       // n, 1-based, LAPACK order
  double* tau;         // kcap
  double* rdiag;       // kcap
  int* info;           // k, nsteps, nblocks, status
  int* kbtrace;
  int kbcap;
};

__device__ __forceinline__ bool qb_barrier(unsigned* bar, unsigned target, int* s_fail) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(bar, 1u);
    unsigned v, spins = 0;
    do {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(bar) : "memory");
    } while (v < target && ++spins < QB_SPIN);
    if (v < target) *s_fail = 1;            // never hang the box: every CTA times out on its own and reports
    __threadfence();
  }
  __syncthreads();
  return *s_fail == 0;
}

__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(d0), "+d"(d1)
               : "d"(a), "d"(b));
}

__device__ __forceinline__ double block_sum(double x, double* red) {      // all threads; red: [QB_WARPS + 1] doubles
  x = warp_sum(x);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = x;
  __syncthreads();
  if (threadIdx.x < 32) {
    double y = threadIdx.x < QB_WARPS ? red[threadIdx.x] : 0.0;
    y = warp_sum(y);
    if (threadIdx.x == 0) red[QB_WARPS] = y;
  }
  __syncthreads();
  return red[QB_WARPS];
}

__global__ void __launch_bounds__(QB_THREADS, 1) qrcp_blocked_kernel(QbParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int G = gridDim.x, cta = blockIdx.x;
  const int64_t m = p.m, lda = p.lda;
  const int64_t col0 = (int64_t)cta * p.cpc;
  const int ncols = (int)max((int64_t)0, min((int64_t)p.cpc, p.n - col0));
  // ---- shared memory ----
  double* Fs = reinterpret_cast<double*>(smem_raw);           // [QB_CPC][QB_FLD] rows of F of my columns
  double* vch = Fs + QB_CPC * QB_FLD;                          // [QB_VCH]         chunk of the Householder vector
  double* svn1 = vch + QB_VCH;                                // [QB_CPC]
  double* svn2 = svn1 + QB_CPC;                               // [QB_CPC]
  double* saux = svn2 + QB_CPC;                               // [QB_NB]  aux_i = -tau V_i' v
  double* sprow = saux + QB_NB;                               // [QB_NB]  pivot-row entries of the block's reflectors
  double* red = sprow + QB_NB;                                // [QB_WARPS + 1 + 3]
  int* slpos = reinterpret_cast<int*>(red + QB_WARPS + 4);    // [QB_CPC] logical position (< s: pivoted)
  int* sflag = slpos + QB_CPC;                                // [QB_CPC] flagged in the current block
  int* sbc = sflag + QB_CPC;                                  // [QB_NB]  physical columns of the block's steps
  int* slive = sbc + QB_NB;                                   // [QB_CPC] compact list of live local columns (block end)
  __shared__ int s_fail, s_w[4], s_nlive, s_anyflag;
  __shared__ double s_wkey;
  __shared__ QbCand s_wc[QB_WARPS];
  if (tid == 0) s_fail = 0;
  unsigned epoch = 0;

  // ---- prologue: column norms (src/pqr.jl:376-385), exact power-of-two scaling ----
  for (int lc = warp; lc < ncols; lc += QB_WARPS) {
    const double* g = p.A + (col0 + lc) * lda;
    double amax = 0.0;
    for (int64_t r = lane; r < m; r += 32) amax = fmax(amax, fabs(g[r]));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) amax = fmax(amax, __shfl_xor_sync(0xffffffffu, amax, o));
    const int e = amax > 0.0 ? ilogb(amax) : 0;
    const double sc = scalbn(1.0, -e);
    double ss = 0.0;
    for (int64_t r = lane; r < m; r += 32) {
      const double x = g[r] * sc;
      ss = fma(x, x, ss);
    }
    ss = warp_sum(ss);
    if (lane == 0) {
      const double nrm = amax > 0.0 ? scalbn(sqrt(ss), e) : 0.0;
      svn1[lc] = nrm;
      svn2[lc] = nrm;
      slpos[lc] = (int)(col0 + lc);
      sflag[lc] = 0;
    }
  }
  __syncthreads();

  // local candidate: argmax of vn1 over my live columns (lpos >= s); ties -> smaller logical position
  auto publish = [&](int s, int par, int flag) {
    double key = -1.0;
    int lp = 0x7fffffff, ph = -1;
    for (int lc = tid; lc < ncols; lc += QB_THREADS) {
      const int q = slpos[lc];
      if (q >= s) {
        const double v = svn1[lc];
        if (cand_better(v, q, key, lp)) {
          key = v;
          lp = q;
          ph = (int)(col0 + lc);
        }
      }
    }
    const int wl = warp_argmax(key, lp);
    if (lane == wl) {
      QbCand c;
      c.key = key;
      c.lp = lp;
      c.phys = ph;
      c.flag = flag;
      c.pad[0] = c.pad[1] = c.pad[2] = 0;
      s_wc[warp] = c;
    }
    __syncthreads();
    if (warp == 0) {
      QbCand c = lane < QB_WARPS ? s_wc[lane] : QbCand{-1.0, 0x7fffffff, -1, 0, {0, 0, 0}};
      const int w2 = warp_argmax(c.key, c.lp);
      if (lane == w2) p.cand[(size_t)par * G + cta] = c;
    }
  };

  // candidates of parity `par` -> (physical column, logical position, key, any flag); every CTA sees the same G entries
  auto select = [&](int par) {
    if (warp == 0) {
      double key = -1.0;
      int lp = 0x7fffffff, ph = -1, fl = 0;
      for (int g = lane; g < G; g += 32) {
        const int4* cp = reinterpret_cast<const int4*>(p.cand + (size_t)par * G + g);
        const int4 c0 = __ldcg(cp), c1 = __ldcg(cp + 1);
        const double ck = __hiloint2double(c0.y, c0.x);
        fl |= c1.x;
        if (cand_better(ck, c0.z, key, lp)) {
          key = ck;
          lp = c0.z;
          ph = c0.w;
        }
      }
      const int wl = warp_argmax(key, lp);
      const int wph = __shfl_sync(0xffffffffu, ph, wl), wlp = __shfl_sync(0xffffffffu, lp, wl);
      const double wkey = __shfl_sync(0xffffffffu, key, wl);
      const int anyf = __any_sync(0xffffffffu, fl != 0) ? 1 : 0;
      if (lane == 0) {
        s_w[0] = wkey >= 0.0 ? wph : -1;
        s_w[1] = wlp;
        s_anyflag = anyf;
        s_wkey = wkey;
      }
    }
    __syncthreads();
  };

  // dlaqps' closing dgemm on the FP64 tensor cores: A[s:, live] -= V[s:, 0:kb] F[live, 0:kb]'.
  // A warp takes 8-row stripes; per stripe the V fragments (negated) are loaded once and swept over the CTA's live
  // columns in groups of 8: D(8 rows x 8 cols) = (-V)(8 x 4) F'(4 x 8) + C, kb/4 DMMAs per tile.
  auto trailing_update = [&](int s, int kb) {
    const int kb4 = (kb + 3) & ~3;
    if (tid == 0) {
      int nl = 0;
      for (int lc = 0; lc < ncols; ++lc)
        if (slpos[lc] >= s) slive[nl++] = lc;
      s_nlive = nl;
      while (nl & 7) slive[nl++] = -1;
    }
    for (int e = tid; e < ncols * QB_NB; e += QB_THREADS)
      if ((e & (QB_NB - 1)) >= kb) Fs[(e >> 5) * QB_FLD + (e & (QB_NB - 1))] = 0.0;
    __syncthreads();
    const int nl = s_nlive, ngrp = (nl + 7) >> 3;
    const int row = lane >> 2, q = lane & 3;
    const bool vec16 = ((lda & 1) == 0) && ((reinterpret_cast<uintptr_t>(p.A) & 15) == 0);
    if (vec16) {
      // TRANSPOSED tiles: MMA rows = 8 of my columns, MMA columns = 8 matrix rows, so a lane's two accumulators are two
      // CONSECUTIVE ROWS of one column -- one 16-byte load / store.  D(j, r) = F(j, 0:kb) (-V(r, 0:kb))' + C(j, r).
      // A warp takes 16-row stripes (two row tiles) and sweeps the column groups two at a time: four 16-byte loads in
      // flight per lane.  The region starts on an even row; V is masked to zero above row s.
      const int64_t sb = s & ~int64_t(1);
      const int64_t nst = (m - sb + 15) >> 4;
      for (int64_t st = warp; st < nst && nl > 0; st += QB_WARPS) {
        const int64_t rb = sb + 16 * st;
        double vb[2][QB_NB / 4];                       // B fragments: B[k = i][n = row] = -V[rb + 8t + row][i]
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          const int64_t r = rb + 8 * t + row;
#pragma unroll
          for (int kq = 0; kq < QB_NB / 4; ++kq) {
            const int i = 4 * kq + q;
            vb[t][kq] = (i < kb && r >= s && r < m) ? -__ldcg(p.A + r + (int64_t)sbc[i] * lda) : 0.0;
          }
        }
        for (int g = 0; g < ngrp; g += 2) {
          double2 c[2][2];
          double2* cp[2][2];
          int lca[2];
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            lca[u] = (g + u < ngrp) ? slive[8 * (g + u) + row] : -1;          // A fragment row / C row: column j = lane / 4
#pragma unroll
            for (int t = 0; t < 2; ++t) {
              const int64_t r = rb + 8 * t + 2 * q;                            // C columns 2q, 2q + 1 -> rows r, r + 1
              cp[u][t] = (lca[u] >= 0 && r < m) ? reinterpret_cast<double2*>(p.A + (col0 + lca[u]) * lda + r) : nullptr;
            }
          }
#pragma unroll
          for (int u = 0; u < 2; ++u)
#pragma unroll
            for (int t = 0; t < 2; ++t) {
              c[u][t] = make_double2(0.0, 0.0);
              if (cp[u][t]) {
                const int64_t r = rb + 8 * t + 2 * q;
                if (r + 1 < m) c[u][t] = *cp[u][t];
                else c[u][t].x = *reinterpret_cast<double*>(cp[u][t]);         // odd last row of the matrix
              }
            }
#pragma unroll
          for (int kq = 0; kq < QB_NB / 4; ++kq)
            if (kq * 4 < kb4) {
#pragma unroll
              for (int u = 0; u < 2; ++u) {
                const double fa = lca[u] >= 0 ? Fs[lca[u] * QB_FLD + 4 * kq + q] : 0.0;
#pragma unroll
                for (int t = 0; t < 2; ++t) dmma884(c[u][t].x, c[u][t].y, fa, vb[t][kq]);
              }
            }
#pragma unroll
          for (int u = 0; u < 2; ++u)
#pragma unroll
            for (int t = 0; t < 2; ++t)
              if (cp[u][t]) {
                const int64_t r = rb + 8 * t + 2 * q;
                if (r + 1 < m) *cp[u][t] = c[u][t];
                else *reinterpret_cast<double*>(cp[u][t]) = c[u][t].x;
              }
        }
      }
      __syncthreads();
      return;
    }
    const int64_t nstripes = (m - s + 7) >> 3;
    for (int64_t st = warp; st < nstripes && nl > 0; st += QB_WARPS) {
      const int64_t r = (int64_t)s + 8 * st + row;
      const bool rin = r < m;
      double va[QB_NB / 4];
#pragma unroll
      for (int kq = 0; kq < QB_NB / 4; ++kq) {
        const int i = 4 * kq + q;
        va[kq] = (i < kb && rin) ? -__ldcg(p.A + r + (int64_t)sbc[i] * lda) : 0.0;
      }
      for (int g = 0; g < ngrp; ++g) {
        const int lcb = slive[8 * g + row];                                       // B fragment: column n = lane / 4
        const int lc0 = slive[8 * g + 2 * q], lc1 = slive[8 * g + 2 * q + 1];    // C fragment: columns 2q, 2q + 1
        double* c0p = (lc0 >= 0 && rin) ? p.A + (col0 + lc0) * lda + r : nullptr;
        double* c1p = (lc1 >= 0 && rin) ? p.A + (col0 + lc1) * lda + r : nullptr;
        double d0 = c0p ? *c0p : 0.0;
        double d1 = c1p ? *c1p : 0.0;
#pragma unroll
        for (int kq = 0; kq < QB_NB / 4; ++kq)
          if (kq * 4 < kb4) {
            const double fb = lcb >= 0 ? Fs[lcb * QB_FLD + 4 * kq + q] : 0.0;
            dmma884(d0, d1, va[kq], fb);
          }
        if (c0p) *c0p = d0;
        if (c1p) *c1p = d1;
      }
    }
    __syncthreads();
  };

  const int lastrk = (int)min(m, p.n);
  const int smax = min(p.kcap, lastrk);        // pivot steps never exceed this
  int s = 0, jblk = 0, cnt = 0, jb = min(p.nb, p.kcap), nblocks = 0, kres = -1;
  double ptol = 0.0;
  bool failed = false;
  publish(0, 0, 0);
  if (!qb_barrier(p.bar, ++epoch * G, &s_fail)) failed = true;

  long long tph[4] = {0, 0, 0, 0}, tlast = clock64();
#define QB_TICK(i) { const long long _t = clock64(); tph[i] += _t - tlast; tlast = _t; }
  while (!failed) {
    const int par = s & 1;
    select(par);
    if (s == 0) ptol = fmax(p.atol, p.rtol * fmax(s_wkey, 0.0));      // src/pqr.jl:386-389 (step-0 keys are the norms)

    // ---- block end: a column was flagged in the step just done, the block is full, or no step is left ----
    if (cnt > 0 && (s_anyflag || cnt == jb || s >= smax)) {
      const int kb = cnt;
      const bool renorm = s_anyflag != 0;
      __syncthreads();
      if (s < min((int64_t)p.n, m)) trailing_update(s, kb);
      // flagged columns: norm of the UPDATED column from row s down (dlaqps: the loop over LSTICC)
      for (int lc = warp; lc < ncols; lc += QB_WARPS) {
        if (!sflag[lc] || slpos[lc] < s) continue;
        const double* g = p.A + (col0 + lc) * lda;
        double ss = 0.0;
        for (int64_t r = s + lane; r < m; r += 32) ss = fma(g[r], g[r], ss);
        ss = warp_sum(ss);
        if (lane == 0) {
          const double nn = sqrt(ss);
          svn1[lc] = nn;
          svn2[lc] = nn;
        }
      }
      __syncthreads();
      for (int lc = tid; lc < ncols; lc += QB_THREADS) sflag[lc] = 0;
      if (cta == 0 && tid == 0 && nblocks < p.kbcap) p.kbtrace[nblocks] = kb;
      ++nblocks;
      // rank test on the diagonal of this block (src/pqr.jl:409-414)
      if (fabs(__ldcg(p.rdiag + s - 1)) <= ptol) {
        for (int i = jblk; i < s; ++i)
          if (fabs(__ldcg(p.rdiag + i)) <= ptol) {
            kres = i;
            break;
          }
      }
      if (kres < 0 && s >= smax) kres = smax;
      if (kres >= 0) break;
      jblk = s;
      cnt = 0;
      jb = min(p.nb, p.kcap - jblk);
      if (renorm) {
        // the norms of the flagged columns changed: fresh candidates for step s, one more exchange
        publish(s, par ^ 1, 0);
        if (!qb_barrier(p.bar, ++epoch * G, &s_fail)) {
          failed = true;
          break;
        }
        select(par ^ 1);
      }
    }
    if (s >= smax) {          // (only reachable with cnt == 0: nothing was factored, e.g. kcap == 0)
      kres = smax;
      break;
    }
    QB_TICK(0)                                    // winner selection + block ends
    const int pw = s_w[0], pwlp = s_w[1];       // pivot of step s: physical column, its logical position
    if (pw < 0) {
      failed = true;
      break;
    }
    __syncthreads();

    // ---- ownership: the column at logical position s moves to the winner's position; the winner takes s ----
    for (int lc = tid; lc < ncols; lc += QB_THREADS) {
      const int ph = (int)(col0 + lc);
      if (ph == pw) slpos[lc] = s;
      else if (slpos[lc] == s) slpos[lc] = pwlp;
    }
    if (tid == 0) sbc[cnt] = pw;
    __syncthreads();

    // ---- owner of the pivot column: bring it up to date, dlarfg, aux ----
    if (pw >= col0 && pw < col0 + ncols) {
      const int lcw = (int)(pw - col0);
      double* aw = p.A + (int64_t)pw * lda;
      const double* Fw = Fs + lcw * QB_FLD;
      // (a) a <- a - V[s:, 0:cnt] Fw[0:cnt]   (dlaqps: "apply previous Householder reflectors to column K")
      double ss = 0.0, alpha = 0.0;
      for (int64_t r = s + tid; r < m; r += QB_THREADS) {
        double a = aw[r];
        int i = 0;
        for (; i + 8 <= cnt; i += 8) {               // eight independent loads in flight per thread
          double x[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) x[u] = __ldcg(p.A + r + (int64_t)sbc[i + u] * lda);
#pragma unroll
          for (int u = 0; u < 8; ++u) a = fma(-x[u], Fw[i + u], a);
        }
        for (; i < cnt; ++i) a = fma(-__ldcg(p.A + r + (int64_t)sbc[i] * lda), Fw[i], a);
        aw[r] = a;
        if (r == s) alpha = a;
        else ss = fma(a, a, ss);
      }
      ss = block_sum(ss, red);
      alpha = block_sum(alpha, red);                    // exactly one thread holds it
      double beta, tau, scale;
      if (s >= m - 1 || ss == 0.0) {
        beta = alpha;
        tau = 0.0;
        scale = 0.0;                                    // H = I (dlarfg: xnorm == 0)
      } else {
        beta = -copysign(sqrt(fma(alpha, alpha, ss)), alpha);
        tau = (beta - alpha) / beta;
        scale = 1.0 / (alpha - beta);
      }
      // (b) v = a[s+1:] / (alpha - beta);  the diagonal entry keeps beta
      for (int64_t r = s + 1 + tid; r < m; r += QB_THREADS) aw[r] *= scale;
      if (tid == 0) {
        aw[s] = beta;
        p.hdr[0] = tau;
        p.hdr[1] = beta;
        p.tau[s] = tau;
        p.rdiag[s] = beta;
        p.jpvt[s] = (int64_t)pw + 1;
      }
      __syncthreads();
      // aux_i = -tau (V_i[s:]' v),  v = [1; a[s+1:]]   (dlaqps: AUXV, the incremental update of F)
      for (int i = warp; i < cnt; i += QB_WARPS) {
        const double* vi = p.A + (int64_t)sbc[i] * lda;
        double d = 0.0;
        for (int64_t r = s + 1 + lane; r < m; r += 32) d = fma(__ldcg(vi + r), aw[r], d);
        d = warp_sum(d);
        if (lane == 0) p.hdr[8 + i] = -tau * (__ldcg(vi + s) + d);
      }
    }
    if (!qb_barrier(p.bar, ++epoch * G, &s_fail)) {
      failed = true;
      break;
    }

    QB_TICK(1)                                    // the owner's column update + dlarfg (everyone else waits)
    // ---- every CTA: the new column of F over its live columns, pivot row, norm downdate ----
    const double tau = __ldcg(p.hdr);
    if (tid < cnt) {
      saux[tid] = __ldcg(p.hdr + 8 + tid);
      sprow[tid] = __ldcg(p.A + s + (int64_t)sbc[tid] * lda);
    }
    const double* vg = p.A + (int64_t)pw * lda;
    double acc[QB_CPW];
#pragma unroll
    for (int jj = 0; jj < QB_CPW; ++jj) acc[jj] = 0.0;
    // chunks start on an even row so that 16-byte loads stay aligned (rows below s read as zero)
    const bool vec16 = ((lda & 1) == 0) && ((reinterpret_cast<uintptr_t>(p.A) & 15) == 0);
    for (int64_t c0 = s & ~int64_t(1); c0 < m; c0 += QB_VCH) {
      const int64_t c1 = min(m, c0 + QB_VCH);
      __syncthreads();
      for (int64_t r = c0 + tid; r < c0 + QB_VCH; r += QB_THREADS)
        vch[r - c0] = (r == s) ? 1.0 : ((r > s && r < c1) ? __ldcg(vg + r) : 0.0);
      __syncthreads();
#pragma unroll
      for (int jj = 0; jj < QB_CPW; ++jj) {
        const int lc = warp + QB_WARPS * jj;
        if (lc >= ncols || slpos[lc] <= s) continue;            // warp-uniform: pivoted (or the pivot itself)
        const double* g = p.A + (col0 + lc) * lda;
        double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
        if (vec16) {
          // rows [c0, ce) in pairs (ce even); a stale entry above s meets a zero of the vector
          const int64_t ce = c0 + ((c1 - c0) & ~int64_t(1));
          const double2* g2 = reinterpret_cast<const double2*>(g + c0);
          const double2* v2 = reinterpret_cast<const double2*>(vch);
          const int64_t np2 = (ce - c0) >> 1;
          int64_t i = lane;
          for (; i + 96 < np2; i += 128) {
            const double2 x0 = g2[i], x1 = g2[i + 32], x2 = g2[i + 64], x3 = g2[i + 96];
            const double2 y0 = v2[i], y1 = v2[i + 32], y2 = v2[i + 64], y3 = v2[i + 96];
            a0 = fma(x0.x, y0.x, a0);
            a1 = fma(x0.y, y0.y, a1);
            a2 = fma(x1.x, y1.x, a2);
            a3 = fma(x1.y, y1.y, a3);
            a0 = fma(x2.x, y2.x, a0);
            a1 = fma(x2.y, y2.y, a1);
            a2 = fma(x3.x, y3.x, a2);
            a3 = fma(x3.y, y3.y, a3);
          }
          for (; i < np2; i += 32) {
            const double2 x0 = g2[i], y0 = v2[i];
            a0 = fma(x0.x, y0.x, a0);
            a1 = fma(x0.y, y0.y, a1);
          }
          if (lane == 0 && ce < c1) a2 = fma(g[ce], vch[ce - c0], a2);      // odd last row of the matrix
        } else {
          int64_t r = c0 + lane;
          for (; r + 96 < c1; r += 128) {
            a0 = fma(g[r], vch[r - c0], a0);
            a1 = fma(g[r + 32], vch[r + 32 - c0], a1);
            a2 = fma(g[r + 64], vch[r + 64 - c0], a2);
            a3 = fma(g[r + 96], vch[r + 96 - c0], a3);
          }
          for (; r < c1; r += 32) a0 = fma(g[r], vch[r - c0], a0);
        }
        acc[jj] += (a0 + a1) + (a2 + a3);
      }
    }
    const bool downdate = s < lastrk - 1;
    int myflag = 0;
#pragma unroll
    for (int jj = 0; jj < QB_CPW; ++jj) {
      const int lc = warp + QB_WARPS * jj;
      if (lc >= ncols || slpos[lc] <= s) continue;
      const double dot = warp_sum(acc[jj]);
      if (lane == 0) {
        double* Fr = Fs + lc * QB_FLD;
        double* ap = p.A + (col0 + lc) * lda + s;
        double f = tau * dot;
        for (int i = 0; i < cnt; ++i) f = fma(Fr[i], saux[i], f);
        Fr[cnt] = f;
        double a = *ap;
        for (int i = 0; i < cnt; ++i) a = fma(-Fr[i], sprow[i], a);
        a -= f;                                                       // the new reflector has v[s] = 1
        *ap = a;
        // LAWN-176 downdate (dlaqps step 8)
        const double v1 = svn1[lc];
        if (downdate && v1 != 0.0) {
          double t = fabs(a) / v1;
          t = fmax(0.0, (1.0 + t) * (1.0 - t));
          const double rr = v1 / svn2[lc];
          if (t * (rr * rr) <= TOL3Z) {
            sflag[lc] = 1;
            myflag = 1;
          } else {
            svn1[lc] = v1 * sqrt(t);
          }
        }
      }
    }
    const int anyflag = __syncthreads_or(myflag);
    QB_TICK(2)                                    // F column pass over my columns
    ++cnt;
    ++s;
    publish(s, s & 1, anyflag);
    if (!qb_barrier(p.bar, ++epoch * G, &s_fail)) {
      failed = true;
      break;
    }
    QB_TICK(3)                                    // candidate exchange (barrier)
  }
#undef QB_TICK

  // ---- epilogue: jpvt of the columns never pivoted, result ----
  __syncthreads();
  if (s_fail) failed = true;
  for (int lc = tid; lc < ncols; lc += QB_THREADS) {
    const int q = slpos[lc];
    if (q >= s) p.jpvt[q] = col0 + lc + 1;
  }
  if (cta == 0 && tid == 0) {
    p.info[0] = failed ? -1 : kres;
    p.info[1] = s;
    p.info[2] = nblocks;
    p.info[3] = failed ? 1 : 0;
    for (int i = 0; i < 4; ++i) p.info[4 + i] = (int)(tph[i] >> 10);      // kilo-cycles per phase (CTA 0; diagnostic)
    for (int i = 8; i < 12; ++i) p.info[i] = 0;
  }
}

}  // namespace

size_t bra_qrcp_blocked_smem() {
  return (size_t)(QB_CPC * QB_FLD + QB_VCH + 2 * QB_CPC + 2 * QB_NB + QB_WARPS + 4) * 8 + (size_t)(3 * QB_CPC + QB_NB) * 4 + 64;
}

// Can the blocked kernel take this shape?  (tall problems: l rows, n columns)
bool bra_qrcp_blocked_ok(int64_t l, int64_t n, int nb, int num_sms) {
  static const char* ev = getenv("BRA_QRCP_BLOCKED");
  if (ev && atoi(ev) == 0) return false;
  const int G = num_sms < QB_MAXG ? num_sms : QB_MAXG;
  return nb >= 1 && nb <= QB_NB && n >= 1 && (n + G - 1) / G <= QB_CPC && l >= 1 && l < (int64_t(1) << 31) &&
         n < (int64_t(1) << 31);
}

int bra_qrcp_blocked_run(bra_ctx* ctx, double* B, int64_t ldb, int64_t l, int64_t n, int kcap, int nb, double atol,
                         double rtol, QrcpOut* out) {
  const int nbe = nb < kcap ? nb : kcap;
  int G = ctx->num_sms < QB_MAXG ? ctx->num_sms : QB_MAXG;
  const int64_t maxG = (n + 7) / 8;
  if (maxG < G) G = (int)(maxG < 1 ? 1 : maxG);
  const int cpc = (int)((n + G - 1) / G);
  BRA_CUDA(ctx->jpvt.reserve((size_t)n * 8));
  BRA_CUDA(ctx->tau.reserve((size_t)kcap * 8));
  BRA_CUDA(ctx->rdiag.reserve((size_t)kcap * 8));
  BRA_CUDA(ctx->info.reserve(64));
  BRA_CUDA(ctx->kbtrace.reserve((size_t)(kcap + 1) * 4));
  const size_t cand_bytes = (size_t)2 * QB_MAXG * sizeof(QbCand);
  const size_t work = cand_bytes + (size_t)(8 + QB_NB) * 8 + 256 + (size_t)QB_NB * 4 + 64;
  BRA_CUDA(ctx->fpend.reserve(work));
  BRA_CUDA(cudaMemsetAsync(ctx->fpend.p, 0, work, ctx->stream));
  unsigned char* wp = reinterpret_cast<unsigned char*>(ctx->fpend.p);
  QbParams p;
  p.A = B;
  p.lda = ldb;
  p.m = l;
  p.n = n;
  p.kcap = kcap;
  p.nb = nbe;
  p.atol = atol;
  p.rtol = rtol;
  p.cpc = cpc;
  p.cand = reinterpret_cast<QbCand*>(wp);
  p.hdr = reinterpret_cast<double*>(wp + cand_bytes);
  p.bar = reinterpret_cast<unsigned*>(wp + cand_bytes + (size_t)(8 + QB_NB) * 8);
  p.jpvt = ctx->jpvt.as<int64_t>();
  p.tau = ctx->tau.as<double>();
  p.rdiag = ctx->rdiag.as<double>();
  p.info = ctx->info.as<int>();
  p.kbtrace = ctx->kbtrace.as<int>();
  p.kbcap = kcap + 1;
  const size_t smem = bra_qrcp_blocked_smem();
  BRA_CUDA(cudaFuncSetAttribute(qrcp_blocked_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  void* args[] = {(void*)&p};
  BRA_CUDA(cudaLaunchCooperativeKernel((void*)qrcp_blocked_kernel, dim3(G), dim3(QB_THREADS), args, smem, ctx->stream));
  ctx->launches++;
  BRA_CUDA(cudaMemcpyAsync(ctx->h_info, ctx->info.p, 48, cudaMemcpyDeviceToHost, ctx->stream));
  BRA_CUDA(cudaStreamSynchronize(ctx->stream));
  out->k = ctx->h_info[0];
  out->nsteps = ctx->h_info[1];
  out->nblocks = ctx->h_info[2];
  out->status = ctx->h_info[3];
  if (out->status != 0 || out->k < 0) {
    ctx->set_error("blocked qrcp kernel: grid barrier timeout");
    return BRA_ERR_INTERNAL;
  }
  return BRA_OK;
}

// Early-terminating Householder QR with column pivoting on a short-wide sketch
// B (l x n), semantics of geqp3_adap_main! + LAPACK dlaqps
// (reference: src/pqr.jl:361-418, src/lapack.jl:117-139; dlaqps itself is external,
//  restated in oracle/lra_oracle.py:dlaqps_restated).
//
// B200 design (not a translation of dlaqps):
//  * ONE persistent cooperative kernel runs the whole chain of pivot steps.
//  * Column-slab ownership: CTA c owns columns [c*cpc, (c+1)*cpc); as many of
//    them as fit are cached in shared memory for the lifetime of the kernel, the
//    rest stay in global memory (L2-resident).  Columns are NEVER physically
//    swapped: each column carries its logical LAPACK position `lpos`, which
//    reproduces idamax's first-maximum tie rule and the vn1/vn2 hand-over on a swap.
//  * Reflectors are applied immediately (rank-1, in registers) instead of being
//    deferred into F and a block-end GEMM.  The block structure of dlaqps is kept
//    as bookkeeping only, because it is observable: a flagged column ends the
//    block, the rank test runs at block ends only, pivoting continues to the block
//    end (src/pqr.jl:397-414).
//  * ONE grid-wide exchange per pivot step: every CTA publishes its best
//    candidate (downdated norm, logical position) TOGETHER with that candidate's
//    current column, as self-validating 8-byte words (payload + step stamp, the
//    "LL" idea of NCCL) so no fence or atomic is on the critical path.  Every CTA
//    then reads the 148 headers, picks the winner with warp-shuffle reductions and
//    computes the Householder vector redundantly (bitwise identical everywhere).
//  * The rank/rtol termination test stays on the device.
#include "common.cuh"
#include <cooperative_groups.h>

namespace {

constexpr int QR_THREADS = 512;
constexpr int QR_WARPS = QR_THREADS / 32;
constexpr int HDR16 = 4;                       // header size in 16-byte units
constexpr uint32_t SPIN_LIMIT = 1u << 24;      // exchange timeout (never hang the box)

struct __align__(16) LL16 {
  uint32_t lo, s0, hi, s1;
};

struct QrcpParams {
  double* B;
  int64_t ldb;
  int l;
  int64_t n;
  int kcap;
  int nb;           // effective block size = min(opts.nb, kcap)
  double atol, rtol;
  int cpc;          // columns per CTA
  int csm;          // of those, cached in shared memory
  int meta_smem;    // vn1/vn2/lpos in shared memory?
  double* vn1g;
  double* vn2g;
  int* lposg;
  LL16* rec;        // [2][G][HDR16 + l]
  uint32_t epoch;
  int64_t* jpvt;    // n, 1-based, LAPACK layout
  double* tau;      // kcap
  double* rdiag;    // kcap
  int* info;        // k, nsteps, nblocks, status
  int* kbtrace;
  int kbcap;
};

__device__ __forceinline__ void ll_store(LL16* p, uint32_t lo, uint32_t hi, uint32_t stamp) {
  asm volatile("st.volatile.global.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(lo), "r"(stamp), "r"(hi),
               "r"(stamp)
               : "memory");
}
__device__ __forceinline__ bool ll_load(const LL16* p, uint32_t stamp, uint32_t& lo, uint32_t& hi) {
  uint32_t s0, s1;
  uint32_t spins = 0;
  do {
    asm volatile("ld.volatile.global.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(lo), "=r"(s0), "=r"(hi), "=r"(s1)
                 : "l"(p)
                 : "memory");
    if (s0 == stamp && s1 == stamp) return true;
  } while (++spins < SPIN_LIMIT);
  return false;
}
__device__ __forceinline__ void ll_store_d(LL16* p, double x, uint32_t stamp) {
  ll_store(p, (uint32_t)__double2loint(x), (uint32_t)__double2hiint(x), stamp);
}

__device__ __forceinline__ double warp_sum(double x) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
  return x;
}

// candidate ordering: larger norm wins; ties go to the smaller logical position
// (idamax returns the FIRST maximum).
__device__ __forceinline__ bool cand_better(double v, int lp, double bv, int blp) {
  return (v > bv) || (v == bv && lp < blp);
}

struct Cand {
  double v;
  int lp;     // logical position
  int id;     // physical column (local scan) or CTA index (gather)
  int ps;     // physical column currently at logical position s (or -1)
  int flag;
};

__device__ __forceinline__ Cand cand_merge(Cand a, const Cand& b) {
  if (cand_better(b.v, b.lp, a.v, a.lp)) {
    a.v = b.v;
    a.lp = b.lp;
    a.id = b.id;
  }
  a.ps = max(a.ps, b.ps);
  a.flag |= b.flag;
  return a;
}
__device__ __forceinline__ Cand cand_warp_reduce(Cand c) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    Cand d;
    d.v = __shfl_xor_sync(0xffffffffu, c.v, o);
    d.lp = __shfl_xor_sync(0xffffffffu, c.lp, o);
    d.id = __shfl_xor_sync(0xffffffffu, c.id, o);
    d.ps = __shfl_xor_sync(0xffffffffu, c.ps, o);
    d.flag = __shfl_xor_sync(0xffffffffu, c.flag, o);
    c = cand_merge(c, d);
  }
  return c;
}

constexpr double TOL3Z = 1.0536712127723509e-08;   // sqrt(2^-53) = sqrt(DLAMCH('Epsilon'))

// NR = number of row-registers per lane (rows s + lane + 32*i); NR == 0: generic two-pass loop.
template <int NR>
__global__ void __launch_bounds__(QR_THREADS, 1) qrcp_kernel(QrcpParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int G = gridDim.x, cta = blockIdx.x;
  const int l = p.l;
  const int64_t col0 = (int64_t)cta * p.cpc;
  const int ncols = (int)max((int64_t)0, min((int64_t)p.cpc, p.n - col0));
  const int csm = min(p.csm, ncols);

  // ---- shared memory carve-up ----
  double* vbuf = reinterpret_cast<double*>(smem_raw);             // l
  double* rdblk = vbuf + l;                                       // nb
  double* red = rdblk + p.nb;                                     // 64 doubles scratch
  Cand* credc = reinterpret_cast<Cand*>(red + 64);                // QR_WARPS candidates
  double* cache = reinterpret_cast<double*>(credc + QR_WARPS + 1);// csm * l
  double* vn1 = p.meta_smem ? cache + (size_t)p.csm * l : p.vn1g + col0;
  double* vn2 = p.meta_smem ? vn1 + p.cpc : p.vn2g + col0;
  int* lpos = p.meta_smem ? reinterpret_cast<int*>(vn2 + p.cpc) : p.lposg + col0;
  __shared__ int s_fail;
  if (tid == 0) s_fail = 0;

  auto colptr = [&](int lc) -> double* {
    return lc < csm ? cache + (size_t)lc * l : p.B + (col0 + lc) * p.ldb;
  };

  // ---- prologue: stage the slab, initial column norms (src/pqr.jl:376-385) ----
  for (int lc = warp; lc < ncols; lc += QR_WARPS) {
    const double* g = p.B + (col0 + lc) * p.ldb;
    double* d = colptr(lc);
    double amax = 0.0;
    for (int r = lane; r < l; r += 32) {
      double x = g[r];
      if (lc < csm) d[r] = x;
      amax = fmax(amax, fabs(x));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) amax = fmax(amax, __shfl_xor_sync(0xffffffffu, amax, o));
    double nrm = 0.0;
    if (amax > 0.0) {
      // exact power-of-two scaling: same rounding as the unscaled sum, no overflow/underflow
      int e = ilogb(amax);
      double sc = scalbn(1.0, -e);
      double ss = 0.0;
      for (int r = lane; r < l; r += 32) {
        double x = d[r] * sc;
        ss = fma(x, x, ss);
      }
      ss = warp_sum(ss);
      nrm = scalbn(sqrt(ss), e);
    }
    if (lane == 0) {
      vn1[lc] = nrm;
      vn2[lc] = nrm;
      lpos[lc] = (int)(col0 + lc);
    }
  }
  __syncthreads();

  const int lastrk = (int)min((int64_t)l, p.n);
  int s = 0;            // current pivot step
  int jblk = 0;         // block start
  int cnt = 0;          // steps done in the current block
  int jb = min(p.nb, p.kcap);
  int nblocks = 0;
  int kres = -1;
  double ptol = 0.0;
  int myflag = 0;       // any local column flagged during the previous step
  bool failed = false;

  while (true) {
    // ---- local scan: best candidate, holder of logical position s, flags ----
    Cand c;
    c.v = -1.0;
    c.lp = 0x7fffffff;
    c.id = -1;
    c.ps = -1;
    c.flag = myflag;
    for (int lc = tid; lc < ncols; lc += QR_THREADS) {
      int lp = lpos[lc];
      if (lp >= s) {
        double v = vn1[lc];
        if (cand_better(v, lp, c.v, c.lp)) {
          c.v = v;
          c.lp = lp;
          c.id = lc;
        }
        if (lp == s) c.ps = (int)(col0 + lc);
      }
    }
    c = cand_warp_reduce(c);
    if (lane == 0) credc[warp] = c;
    __syncthreads();
    c = credc[0];
#pragma unroll
    for (int w = 1; w < QR_WARPS; ++w) c = cand_merge(c, credc[w]);
    __syncthreads();     // credc reused below

    // ---- publish: header + the candidate's current column ----
    const uint32_t stamp = p.epoch + (uint32_t)s;
    LL16* myrec = p.rec + ((size_t)(s & 1) * G + cta) * (HDR16 + l);
    if (c.id >= 0) {
      const double* a = colptr(c.id);
      for (int r = s + tid; r < l; r += QR_THREADS) ll_store_d(myrec + HDR16 + r, a[r], stamp);
    }
    if (tid == 0) {
      ll_store_d(myrec + 0, c.v, stamp);
      ll_store(myrec + 1, (uint32_t)c.lp, (uint32_t)(c.id >= 0 ? (int)(col0 + c.id) : -1), stamp);
      ll_store(myrec + 2, (uint32_t)c.ps, (uint32_t)c.flag, stamp);
    }

    // ---- gather the G headers, pick the winner ----
    Cand w;
    w.v = -2.0;
    w.lp = 0x7fffffff;
    w.id = -1;
    w.ps = -1;
    w.flag = 0;
    int wphys = -1;
    const int gw = (G + 31) / 32;   // warps that hold headers
    if (tid < G) {
      const LL16* r = p.rec + ((size_t)(s & 1) * G + tid) * (HDR16 + l);
      uint32_t a0, a1, b0, b1, c0, c1;
      bool ok = ll_load(r + 0, stamp, a0, a1);
      ok = ok && ll_load(r + 1, stamp, b0, b1);
      ok = ok && ll_load(r + 2, stamp, c0, c1);
      if (!ok) s_fail = 1;
      w.v = __hiloint2double((int)a1, (int)a0);
      w.lp = (int)b0;
      wphys = (int)b1;
      w.id = tid;
      w.ps = (int)c0;
      w.flag = (int)c1;
    }
    if (warp < gw) {
      // carry the winner's physical column through the reduction in `id` afterwards
      Cand wr = cand_warp_reduce(w);
      int src = __ffs(__ballot_sync(0xffffffffu, w.id == wr.id && w.id >= 0)) - 1;
      int wp = __shfl_sync(0xffffffffu, wphys, max(src, 0));
      if (lane == 0) {
        credc[warp] = wr;
        red[warp] = (double)wp;
      }
    }
    __syncthreads();
    w = credc[0];
    wphys = (int)red[0];
    for (int q = 1; q < gw; ++q) {
      Cand o = credc[q];
      if (cand_better(o.v, o.lp, w.v, w.lp)) wphys = (int)red[q];
      w = cand_merge(w, o);
    }
    if (s_fail) {
      failed = true;
      break;
    }
    const int wcta = w.id;           // CTA that owns the winner
    const int lw = w.lp;             // winner's logical position
    const int pw = wphys;            // winner's physical column
    const int ps = w.ps;             // physical column at logical position s

    // ---- block bookkeeping for the previous step (needs the gathered flags) ----
    if (cnt > 0 && w.flag) {
      // a column was flagged during step s-1: dlaqps ended its block there
      if (cta == 0 && tid == 0 && nblocks < p.kbcap) p.kbtrace[nblocks] = cnt;
      ++nblocks;
      const int jn = jblk + cnt;
      if (fabs(rdblk[cnt - 1]) <= ptol) {
        for (int i = 0; i < cnt; ++i)
          if (fabs(rdblk[i]) <= ptol) {
            kres = jblk + i;
            break;
          }
      }
      if (kres >= 0) break;
      jblk = jn;
      cnt = 0;
      jb = min(p.nb, p.kcap - jblk);
      // jblk < kcap here: a block that reaches kcap always ends by count
    }
    if (s == 0) ptol = fmax(p.atol, p.rtol * w.v);       // src/pqr.jl:386-389
    __syncthreads();   // rdblk / credc consumers done

    // ---- Householder vector of the winner column (dlarfg), redundantly per CTA ----
    const LL16* wrec = p.rec + ((size_t)(s & 1) * G + wcta) * (HDR16 + l) + HDR16;
    constexpr int XS = 3;                         // covers l - s <= 3 * QR_THREADS rows in registers
    double xs[XS];
    double ss = 0.0;
    {
      int q = 0;
      for (int r = s + tid; r < l; r += QR_THREADS, ++q) {
        uint32_t lo, hi;
        if (!ll_load(wrec + r, stamp, lo, hi)) s_fail = 1;
        double x = __hiloint2double((int)hi, (int)lo);
        if (q < XS) xs[q] = x;
        else vbuf[r - s] = x;                           // l > 1056+: spill through smem
        if (r > s) ss = fma(x, x, ss);
      }
    }
    ss = warp_sum(ss);
    if (lane == 0) red[32 + warp] = ss;
    if (tid == 0) red[63] = xs[0];                        // alpha (row s is always thread 0's first)
    __syncthreads();
    if (s_fail) {
      failed = true;
      break;
    }
    double ssq = 0.0;
#pragma unroll
    for (int q = 0; q < QR_WARPS; ++q) ssq += red[32 + q];
    const double alpha = red[63];
    double beta, tau, scale;
    if (s >= l - 1 || ssq == 0.0) {
      beta = alpha;
      tau = 0.0;
      scale = 0.0;        // v = e_1
    } else {
      beta = -copysign(sqrt(fma(alpha, alpha, ssq)), alpha);
      tau = (beta - alpha) / beta;
      scale = 1.0 / (alpha - beta);
    }
    {
      int q = 0;
      for (int r = s + tid; r < l; r += QR_THREADS, ++q) {
        double x = (q < XS) ? xs[q] : vbuf[r - s];
        vbuf[r - s] = (r == s) ? 1.0 : x * scale;
      }
    }
    if (tid == 0) rdblk[cnt] = beta;
    // ---- ownership updates ----
    if (tid == 0) {
      if (ps >= col0 && ps < col0 + ncols && ps != pw) lpos[ps - col0] = lw;   // column K moves to pvt
      if (wcta == cta) {
        lpos[pw - col0] = s;
        p.jpvt[s] = (int64_t)pw + 1;
        p.tau[s] = tau;
        p.rdiag[s] = beta;
      }
    }
    __syncthreads();
    if (wcta == cta) {
      // store R[s,s] and the reflector into the winner column (LAPACK layout)
      double* a = colptr((int)(pw - col0));
      for (int r = s + tid; r < l; r += QR_THREADS) a[r] = (r == s) ? beta : vbuf[r - s];
    }

    // ---- apply H to the local unpivoted columns, downdate their norms ----
    myflag = 0;
    const bool downdate = (s < lastrk - 1);
    if (NR > 0) {
      double vr[NR > 0 ? NR : 1];
#pragma unroll
      for (int i = 0; i < NR; ++i) {
        int r = s + lane + 32 * i;
        vr[i] = (r < l) ? vbuf[r - s] : 0.0;
      }
      for (int lc = warp; lc < ncols; lc += QR_WARPS) {
        if (lpos[lc] <= s) continue;
        double* a = colptr(lc);
        double ar[NR > 0 ? NR : 1];
        double dot = 0.0;
#pragma unroll
        for (int i = 0; i < NR; ++i) {
          int r = s + lane + 32 * i;
          ar[i] = (r < l) ? a[r] : 0.0;
          dot = fma(ar[i], vr[i], dot);
        }
        dot = warp_sum(dot);
        const double f = tau * dot;
        double ss2 = 0.0;
#pragma unroll
        for (int i = 0; i < NR; ++i) {
          int r = s + lane + 32 * i;
          ar[i] = fma(-f, vr[i], ar[i]);
          if (r < l) a[r] = ar[i];
          if (r > s) ss2 = fma(ar[i], ar[i], ss2);      // rows beyond l hold zeros
        }
        if (downdate) {
          const double rsj = __shfl_sync(0xffffffffu, ar[0], 0);
          const double v1 = vn1[lc];
          if (v1 != 0.0) {
            double t = fabs(rsj) / v1;
            t = fmax(0.0, (1.0 + t) * (1.0 - t));
            const double q = v1 / vn2[lc];
            const double t2 = t * (q * q);
            if (t2 <= TOL3Z) {
              // flagged: dlaqps ends the block and recomputes the norm from rows s+1..l-1
              ss2 = warp_sum(ss2);
              const double nn = sqrt(ss2);
              if (lane == 0) {
                vn1[lc] = nn;
                vn2[lc] = nn;
              }
              myflag = 1;
            } else if (lane == 0) {
              vn1[lc] = v1 * sqrt(t);
            }
          }
        }
      }
    } else {
      for (int lc = warp; lc < ncols; lc += QR_WARPS) {
        if (lpos[lc] <= s) continue;
        double* a = colptr(lc);
        double dot = 0.0;
        for (int r = s + lane; r < l; r += 32) dot = fma(a[r], vbuf[r - s], dot);
        dot = warp_sum(dot);
        const double f = tau * dot;
        double ss2 = 0.0, rsj = 0.0;
        for (int r = s + lane; r < l; r += 32) {
          double x = fma(-f, vbuf[r - s], a[r]);
          a[r] = x;
          if (r > s) ss2 = fma(x, x, ss2);
          else rsj = x;
        }
        if (downdate) {
          rsj = __shfl_sync(0xffffffffu, rsj, 0);
          const double v1 = vn1[lc];
          if (v1 != 0.0) {
            double t = fabs(rsj) / v1;
            t = fmax(0.0, (1.0 + t) * (1.0 - t));
            const double q = v1 / vn2[lc];
            const double t2 = t * (q * q);
            if (t2 <= TOL3Z) {
              ss2 = warp_sum(ss2);
              const double nn = sqrt(ss2);
              if (lane == 0) {
                vn1[lc] = nn;
                vn2[lc] = nn;
              }
              myflag = 1;
            } else if (lane == 0) {
              vn1[lc] = v1 * sqrt(t);
            }
          }
        }
      }
    }
    myflag = __syncthreads_or(myflag);

    // ---- end of step ----
    ++cnt;
    ++s;
    if (cnt == jb) {
      // block ends by count; flags raised in this step are irrelevant (cnt is reset)
      if (cta == 0 && tid == 0 && nblocks < p.kbcap) p.kbtrace[nblocks] = cnt;
      ++nblocks;
      const int jn = jblk + cnt;
      if (fabs(rdblk[cnt - 1]) <= ptol) {
        for (int i = 0; i < cnt; ++i)
          if (fabs(rdblk[i]) <= ptol) {
            kres = jblk + i;
            break;
          }
      }
      if (kres >= 0) break;
      jblk = jn;
      cnt = 0;
      if (jblk >= p.kcap) {
        kres = p.kcap;
        break;
      }
      jb = min(p.nb, p.kcap - jblk);
      myflag = 0;
    }
  }

  // ---- epilogue: write the cached slab back, finish jpvt, report ----
  __syncthreads();
  const int nsteps = s;     // pivoted columns (when we broke out after a gather, step s was not executed)
  for (int lc = warp; lc < csm; lc += QR_WARPS) {
    double* g = p.B + (col0 + lc) * p.ldb;
    const double* d = cache + (size_t)lc * l;
    for (int r = lane; r < l; r += 32) g[r] = d[r];
  }
  for (int lc = tid; lc < ncols; lc += QR_THREADS) {
    int lp = lpos[lc];
    if (lp >= nsteps) p.jpvt[lp] = col0 + lc + 1;
  }
  if (cta == 0 && tid == 0) {
    p.info[0] = failed ? -1 : kres;
    p.info[1] = nsteps;
    p.info[2] = nblocks;
    p.info[3] = failed ? 1 : 0;
  }
}

__global__ void permute_cols_kernel(const double* __restrict__ src, int64_t lds, double* __restrict__ dst,
                                    int64_t ldd, int64_t rows, int64_t n, const int64_t* __restrict__ jpvt1) {
  // dst[:, j] = src[:, jpvt[j]-1]; one warp-group per column, coalesced along rows
  for (int64_t j = blockIdx.x; j < n; j += gridDim.x) {
    const double* s = src + (jpvt1[j] - 1) * lds;
    double* d = dst + j * ldd;
    for (int64_t r = threadIdx.x; r < rows; r += blockDim.x) d[r] = s[r];
  }
}

__global__ void gather_R_kernel(const double* __restrict__ B, int64_t ldb, int64_t n, int k,
                                const int64_t* __restrict__ jpvt1, double* __restrict__ R11,
                                double* __restrict__ R12) {
  // R = triu(B[0:k, p]) split into R11 (k x k, ld k) and R12 (k x (n-k), ld k)  (src/pqr.jl:428)
  for (int64_t j = blockIdx.x; j < n; j += gridDim.x) {
    const double* s = B + (jpvt1[j] - 1) * ldb;
    if (j < k) {
      double* d = R11 + j * (int64_t)k;
      for (int r = threadIdx.x; r < k; r += blockDim.x) d[r] = (r <= j) ? s[r] : 0.0;
    } else {
      double* d = R12 + (j - k) * (int64_t)k;
      for (int r = threadIdx.x; r < k; r += blockDim.x) d[r] = s[r];
    }
  }
}

template <int NR>
cudaError_t launch_qrcp(const QrcpParams& p, int G, size_t smem, cudaStream_t st) {
  cudaError_t e = cudaFuncSetAttribute(qrcp_kernel<NR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  void* args[] = {(void*)&p};
  return cudaLaunchCooperativeKernel((void*)qrcp_kernel<NR>, dim3(G), dim3(QR_THREADS), args, smem, st);
}

}  // namespace

int bra_qrcp_run(bra_ctx* ctx, double* B, int64_t ldb, int l, int64_t n, int kcap, int nb, double atol,
                 double rtol, QrcpOut* out) {
  out->k = 0;
  out->nsteps = 0;
  out->nblocks = 0;
  out->status = 0;
  if (kcap <= 0 || n <= 0 || l <= 0) return BRA_OK;
  const int nbe = nb < kcap ? nb : kcap;

  // grid: one CTA per SM, but keep >= 8 columns per CTA
  int G = ctx->num_sms;
  int64_t maxG = (n + 7) / 8;
  if (maxG < G) G = (int)(maxG < 1 ? 1 : maxG);
  const int cpc = (int)((n + G - 1) / G);

  // shared-memory budget
  const size_t fixed = ((size_t)l + nbe + 64) * 8 + (QR_WARPS + 2) * sizeof(Cand) + 64;
  const size_t budget = (size_t)ctx->smem_optin - 1024;
  const size_t meta = (size_t)cpc * 20 + 16;
  int meta_smem = (fixed + meta <= budget / 2) ? 1 : 0;
  size_t avail = budget - fixed - (meta_smem ? meta : 0);
  int csm = (int)(avail / ((size_t)l * 8));
  if (csm > cpc) csm = cpc;
  const size_t smem = fixed + (meta_smem ? meta : 0) + (size_t)csm * l * 8;

  BRA_CUDA(ctx->vn1.reserve((size_t)n * 8));
  BRA_CUDA(ctx->vn2.reserve((size_t)n * 8));
  BRA_CUDA(ctx->lpos.reserve((size_t)n * 4));
  const size_t rec_bytes = (size_t)2 * G * (HDR16 + l) * sizeof(LL16);
  if (ctx->rec.cap < rec_bytes || ctx->rec_zeroed < rec_bytes || ctx->rec_epoch > 0xF0000000u) {
    BRA_CUDA(ctx->rec.reserve(rec_bytes));
    BRA_CUDA(cudaMemsetAsync(ctx->rec.p, 0, ctx->rec.cap, ctx->stream));
    ctx->rec_zeroed = ctx->rec.cap;
    ctx->rec_epoch = 1;
  }
  BRA_CUDA(ctx->jpvt.reserve((size_t)n * 8));
  BRA_CUDA(ctx->tau.reserve((size_t)kcap * 8));
  BRA_CUDA(ctx->rdiag.reserve((size_t)kcap * 8));
  BRA_CUDA(ctx->info.reserve(64));
  BRA_CUDA(ctx->kbtrace.reserve((size_t)(kcap + 1) * 4));

  QrcpParams p;
  p.B = B;
  p.ldb = ldb;
  p.l = l;
  p.n = n;
  p.kcap = kcap;
  p.nb = nbe;
  p.atol = atol;
  p.rtol = rtol;
  p.cpc = cpc;
  p.csm = csm;
  p.meta_smem = meta_smem;
  p.vn1g = ctx->vn1.as<double>();
  p.vn2g = ctx->vn2.as<double>();
  p.lposg = ctx->lpos.as<int>();
  p.rec = ctx->rec.as<LL16>();
  p.epoch = ctx->rec_epoch;
  p.jpvt = ctx->jpvt.as<int64_t>();
  p.tau = ctx->tau.as<double>();
  p.rdiag = ctx->rdiag.as<double>();
  p.info = ctx->info.as<int>();
  p.kbtrace = ctx->kbtrace.as<int>();
  p.kbcap = kcap + 1;
  ctx->rec_epoch += (uint32_t)l + 8;

  cudaError_t e;
  if (l <= 64) e = launch_qrcp<2>(p, G, smem, ctx->stream);
  else if (l <= 96) e = launch_qrcp<3>(p, G, smem, ctx->stream);
  else if (l <= 160) e = launch_qrcp<5>(p, G, smem, ctx->stream);
  else if (l <= 288) e = launch_qrcp<9>(p, G, smem, ctx->stream);
  else if (l <= 544) e = launch_qrcp<17>(p, G, smem, ctx->stream);
  else e = launch_qrcp<0>(p, G, smem, ctx->stream);
  BRA_CUDA(e);
  ctx->launches++;
  BRA_CUDA(cudaMemcpyAsync(ctx->h_info, ctx->info.p, 16, cudaMemcpyDeviceToHost, ctx->stream));
  BRA_CUDA(cudaStreamSynchronize(ctx->stream));      // the only host sync: read back k
  out->k = ctx->h_info[0];
  out->nsteps = ctx->h_info[1];
  out->nblocks = ctx->h_info[2];
  out->status = ctx->h_info[3];
  if (out->status != 0 || out->k < 0) {
    ctx->set_error("qrcp kernel: exchange timeout");
    return BRA_ERR_INTERNAL;
  }
  return BRA_OK;
}

int bra_permute_cols(bra_ctx* ctx, const double* src, int64_t lds, double* dst, int64_t ldd, int64_t rows,
                     int64_t n, const int64_t* jpvt1) {
  if (n <= 0 || rows <= 0) return BRA_OK;
  int grid = (int)(n < 148 * 8 ? n : 148 * 8);
  permute_cols_kernel<<<grid, 256, 0, ctx->stream>>>(src, lds, dst, ldd, rows, n, jpvt1);
  ctx->launches++;
  BRA_CUDA(cudaGetLastError());
  return BRA_OK;
}

int bra_gather_R(bra_ctx* ctx, const double* B, int64_t ldb, int64_t n, int k, const int64_t* jpvt1,
                 double* R11, double* R12) {
  if (n <= 0 || k <= 0) return BRA_OK;
  int grid = (int)(n < 148 * 8 ? n : 148 * 8);
  gather_R_kernel<<<grid, 128, 0, ctx->stream>>>(B, ldb, n, k, jpvt1, R11, R12);
  ctx->launches++;
  BRA_CUDA(cudaGetLastError());
  return BRA_OK;
}

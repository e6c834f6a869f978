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
//  * ONE grid-wide exchange per pivot step, push style: as soon as a CTA knows
//    its best downdated norm it pushes ONE 32-byte word (norm, logical position,
//    flag; a single 256-bit store, 4-byte payload + 4-byte step stamp per 8 bytes,
//    the "LL" idea of NCCL) into EVERY CTA's private inbox: no fence, no atomic,
//    no shared polling hot-spot, one L2 request per (source, destination) pair.
//    While those words are in flight the CTA runs dlarfg on its own candidate
//    (block-wide sum of squares, beta, tau) and publishes the finished Householder
//    vector in its record (speculatively: only the winner's record is read).  Every
//    CTA then reads its inbox, picks the winner with redux.sync-based argmax and
//    copies the winner's vector -- the sqrt/div latency of dlarfg and the column
//    transfer hide behind the header exchange.
//  * Norm downdate + local argmax are fused into the update sweep; the scalar
//    LAWN-176 arithmetic is vectorised across lanes (lane j <-> j-th column of the warp).
//  * The rank/rtol termination test stays on the device.
#include "common.cuh"
#include <cstdlib>
#include <type_traits>
#include "qrcp_common.cuh"
#include "qrcp_exchange.cuh"

namespace {

struct Cand {
  double v;       // downdated norm vn1 (-1: none)
  int lp;         // logical position
  int id;         // local column index
  int ps;         // physical column at the next logical pivot position (or -1)
  int flag;       // any column flagged during the step
};

// NR = row registers per lane (rows s + lane + 32*i); NR == 0: generic two-pass loop.
template <int NR>
__global__ void __launch_bounds__(QR_THREADS, 1) qrcp_kernel(QrcpParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int G = gridDim.x, cta = blockIdx.x;
  const int l = p.l;
  const int64_t col0 = (int64_t)cta * p.cpc;
  const int ncols = (int)max((int64_t)0, min((int64_t)p.cpc, p.n - col0));
  const int csm = min(p.csm, ncols);
  constexpr int CB = (NR == 0 || NR > 9) ? 1 : (NR > 5 ? 2 : 4);     // columns in flight per warp

  // ---- shared memory carve-up ----
  double* vbuf = reinterpret_cast<double*>(smem_raw);             // l
  double* rdblk = vbuf + l;                                       // nb
  double* hv = rdblk + p.nb;                                      // MAXG  header payloads of the gather
  double* sred = hv + MAXG;                                       // QR_WARPS partial sums of squares
  int* hlp = reinterpret_cast<int*>(sred + QR_WARPS);             // MAXG each
  int* hflag = hlp + MAXG;
  Cand* credc = reinterpret_cast<Cand*>(hflag + MAXG);            // [2][QR_WARPS]
  double* cache = reinterpret_cast<double*>(credc + 2 * QR_WARPS);// csm * l
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
  __syncwarp();

  // Per-warp candidate for pivot step `snext` over the warp's columns (lc = warp + 16*j), with the
  // result goes to credc[snext & 1][warp].
  auto warp_candidate = [&](int snext, int wflag) {
    double bv = -1.0;
    int blp = 0x7fffffff, bid = -1, bps = -1;
    for (int j0 = 0; warp + QR_WARPS * j0 < ncols; j0 += 32) {
      const int lc = warp + QR_WARPS * (j0 + lane);
      double v = -1.0;
      int lp = 0x7fffffff;
      if (lc < ncols) {
        const int q = lpos[lc];
        if (q >= snext) {
          v = p.nopivot ? (q == snext ? 1.0 : -1.0) : vn1[lc];
          lp = q;
          if (q == snext) bps = (int)(col0 + lc);
        }
      }
      const int wl = warp_argmax(v, lp);
      v = __shfl_sync(0xffffffffu, v, wl);
      lp = __shfl_sync(0xffffffffu, lp, wl);
      const int wlc = warp + QR_WARPS * (j0 + wl);
      if (cand_better(v, lp, bv, blp)) {
        bv = v;
        blp = lp;
        bid = (v >= 0.0) ? wlc : -1;
      }
    }
    bps = __reduce_max_sync(0xffffffffu, bps);
    if (lane == 0) {
      Cand c;
      c.v = bv;
      c.lp = blp;
      c.id = bid;
      c.ps = bps;
      c.flag = wflag;
      credc[(snext & 1) * QR_WARPS + warp] = c;
    }
  };

  const int lastrk = (int)min((int64_t)l, p.n);
  int s = 0;            // current pivot step
  int jblk = 0;         // block start
  int cnt = 0;          // steps done in the current block
  int jb = min(p.nb, p.kcap);
  int nblocks = 0;
  int kres = -1;
  double ptol = 0.0;
  bool failed = false;
  __shared__ long long s_tph[6];            // per-phase cycle totals (thread 0 of CTA 0 reports them)
  if (tid == 0) {
    for (int i = 0; i < 5; ++i) s_tph[i] = 0;
    s_tph[5] = clock64();
  }
#define QR_TICK(i) if (tid == 0) { long long _t = clock64(); s_tph[i] += _t - s_tph[5]; s_tph[5] = _t; }

  warp_candidate(0, 0);
  __syncthreads();

  while (true) {
    // ---- merge the warp candidates -> this CTA's candidate for step s ----
    // (every warp does it redundantly: lane w < QR_WARPS holds warp w's candidate)
    Cand c;
    {
      const Cand* cw = credc + (s & 1) * QR_WARPS;
      double v = -1.0;
      int lp = 0x7fffffff, psx = -1, fl = 0;
      if (lane < QR_WARPS) {
        v = cw[lane].v;
        lp = cw[lane].lp;
        psx = cw[lane].ps;
        fl = cw[lane].flag;
      }
      const int wl = warp_argmax(v, lp);
      c = cw[wl];
      c.ps = __reduce_max_sync(0xffffffffu, psx);
      c.flag = (int)__reduce_or_sync(0xffffffffu, (unsigned)fl);
    }
    QR_TICK(0)

    // ---- publish 1: one 32-byte header word into every CTA's inbox ----
    const uint32_t stamp = p.epoch + (uint32_t)s;
    const int par = s & 1;
    const int my_ps = c.ps;              // physical column sitting at logical position s, if this CTA owns it
    const int cand_lc = c.id;
    if (tid < G) ll32_store(p.inbox + ((size_t)par * G + tid) * G + cta, c.v, c.lp, c.flag, stamp);

    // ---- publish 2 (while the headers fly): dlarfg on my candidate, finished vector into my record ----
    LL16* myrec = p.rec + ((size_t)par * G + cta) * (l + RECH);
    {
      const double* a = (cand_lc >= 0) ? colptr(cand_lc) : vbuf;
      double ss = 0.0;
      double xr[2] = {0.0, 0.0};          // my rows s+1+tid, s+1+tid+512 of the candidate
      if (cand_lc >= 0) {
        const int r0 = s + 1 + tid, r1 = r0 + QR_THREADS;
        if (r0 < l) xr[0] = a[r0];
        if (r1 < l) xr[1] = a[r1];
        ss = fma(xr[0], xr[0], xr[1] * xr[1]);
        for (int r = r1 + QR_THREADS; r < l; r += QR_THREADS) ss = fma(a[r], a[r], ss);
      }
      ss = warp_sum(ss);
      if (lane == 0) sred[warp] = ss;
      __syncthreads();
      // only warps that still hold rows of the candidate (and warp 0, which publishes tau and beta) run the FP64
      // square root and divisions: the FP64 pipe issues per warp, not per active lane
      if (cand_lc >= 0 && (warp == 0 || s + 1 + (warp << 5) < l)) {
        // pairwise (4 levels instead of a 16-long dependent chain of FP64 adds)
        double t8[8];
#pragma unroll
        for (int w = 0; w < 8; ++w) t8[w] = sred[2 * w] + sred[2 * w + 1];
        const double ssq = ((t8[0] + t8[1]) + (t8[2] + t8[3])) + ((t8[4] + t8[5]) + (t8[6] + t8[7]));
        const double alpha = a[s];
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
        const int r0 = s + 1 + tid, r1 = r0 + QR_THREADS;
        if (r0 < l) ll_store_d(myrec + RECH + r0, xr[0] * scale, stamp);
        if (r1 < l) ll_store_d(myrec + RECH + r1, xr[1] * scale, stamp);
        for (int r = r1 + QR_THREADS; r < l; r += QR_THREADS) ll_store_d(myrec + RECH + r, a[r] * scale, stamp);
        if (tid == 0) {
          ll_store_d(myrec + 0, tau, stamp);
          ll_store_d(myrec + 1, beta, stamp);
          ll_store(myrec + 2, (uint32_t)(col0 + cand_lc), 0u, stamp);
        }
      }
    }
    QR_TICK(1)

    // ---- gather my inbox (G contiguous 32-byte words), pick the winner ----
    // two levels: the ceil(G/32) warps that hold the headers in registers reduce them with redux.sync right away and
    // park one partial winner each; after the barrier every warp merges those <= 5 partials (instead of every warp
    // scanning all G headers out of shared memory)
    const int GW = (G + 31) >> 5;
    if (warp < GW) {
      double v = -1.0;
      int lp = 0x7fffffff, fl = 0;
      if (tid < G && !ll32_load(p.inbox + ((size_t)par * G + cta) * G + tid, stamp, v, lp, fl)) s_fail = 1;
      const int wl = warp_argmax(v, lp);
      const double bv = __shfl_sync(0xffffffffu, v, wl);
      const int blp = __shfl_sync(0xffffffffu, lp, wl);
      const unsigned af = __reduce_or_sync(0xffffffffu, (unsigned)fl);
      if (lane == 0) {
        hv[warp] = bv;
        hlp[warp] = blp;
        hflag[warp] = (int)af;
        hflag[MAXG / 2 + warp] = (warp << 5) + wl;        // source CTA of this partial winner
      }
    }
    __syncthreads();
    if (s_fail) {
      failed = true;
      break;
    }
    int wcta;
    {
      const double v = (lane < GW) ? hv[lane] : -1.0;
      const int lp = (lane < GW) ? hlp[lane] : 0x7fffffff;
      const int fl = (lane < GW) ? hflag[lane] : 0;
      const int src = (lane < GW) ? hflag[MAXG / 2 + lane] : -1;
      const int wl = warp_argmax(v, lp);
      wcta = __shfl_sync(0xffffffffu, src, wl);
      c.v = __shfl_sync(0xffffffffu, v, wl);        // reuse c as the gathered result
      c.lp = __shfl_sync(0xffffffffu, lp, wl);
      c.flag = (int)__reduce_or_sync(0xffffffffu, (unsigned)fl);
    }
    const int lw = c.lp;               // winner's logical position
    QR_TICK(2)

    // ---- block bookkeeping for the previous step (needs the gathered flags) ----
    if (cnt > 0 && c.flag) {
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
    if (s == 0) ptol = fmax(p.atol, p.rtol * c.v);       // src/pqr.jl:386-389

    // ---- the winner's Householder vector (already scaled by its owner), tau, beta, physical column ----
    const LL16* wrec = p.rec + ((size_t)par * G + wcta) * (l + RECH);
    double tau, beta;
    int pw;
    {
      // all loads of this thread are issued before the first stamp check: one L2 round trip, not five
      const int r0 = s + 1 + tid, r1 = r0 + QR_THREADS;
      uint32_t q[5][4];
      ll_ld(wrec + 0, q[0][0], q[0][1], q[0][2], q[0][3]);
      ll_ld(wrec + 1, q[1][0], q[1][1], q[1][2], q[1][3]);
      ll_ld(wrec + 2, q[2][0], q[2][1], q[2][2], q[2][3]);
      if (r0 < l) ll_ld(wrec + RECH + r0, q[3][0], q[3][1], q[3][2], q[3][3]);
      if (r1 < l) ll_ld(wrec + RECH + r1, q[4][0], q[4][1], q[4][2], q[4][3]);
#pragma unroll
      for (int i = 0; i < 5; ++i) {
        const bool need = (i < 3) || (i == 3 && r0 < l) || (i == 4 && r1 < l);
        if (need && (q[i][1] != stamp || q[i][3] != stamp)) {
          const LL16* src = (i < 3) ? wrec + i : wrec + RECH + (i == 3 ? r0 : r1);
          if (!ll_load(src, stamp, q[i][0], q[i][2])) s_fail = 1;
        }
      }
      tau = __hiloint2double((int)q[0][2], (int)q[0][0]);
      beta = __hiloint2double((int)q[1][2], (int)q[1][0]);
      pw = (int)q[2][0];
      if (tid == 0) vbuf[0] = 1.0;
      if (r0 < l) vbuf[r0 - s] = __hiloint2double((int)q[3][2], (int)q[3][0]);
      if (r1 < l) vbuf[r1 - s] = __hiloint2double((int)q[4][2], (int)q[4][0]);
      for (int r = r1 + QR_THREADS; r < l; r += QR_THREADS) {
        uint32_t lo, hi;
        if (!ll_load(wrec + RECH + r, stamp, lo, hi)) s_fail = 1;
        vbuf[r - s] = __hiloint2double((int)hi, (int)lo);
      }
    }
    if (tid == 0) {
      rdblk[cnt] = beta;
      // ---- ownership updates ----
      if (my_ps >= 0 && my_ps != pw) lpos[my_ps - col0] = lw;   // column K moves to pvt
      if (wcta == cta) {
        lpos[pw - col0] = s;
        p.jpvt[s] = (int64_t)pw + 1;
        p.tau[s] = tau;
        p.rdiag[s] = beta;
      }
    }
    __syncthreads();
    if (s_fail) {
      failed = true;
      break;
    }
    QR_TICK(3)
    if (wcta == cta) {
      // store R[s,s] and the reflector into the winner column (LAPACK layout)
      double* a = colptr((int)(pw - col0));
      for (int r = s + tid; r < l; r += QR_THREADS) a[r] = (r == s) ? beta : vbuf[r - s];
    }

    // ---- apply H to the warp's unpivoted columns; downdate their norms (vectorised across lanes) ----
    const bool downdate = (s < lastrk - 1) && !p.nopivot;      // no pivoting: the norms are never looked at
    int wflag = 0;
    double vr[NR > 0 ? NR : 1];
    if (NR > 0) {
#pragma unroll
      for (int i = 0; i < NR; ++i) {
        const int r = s + lane + 32 * i;
        vr[i] = (r < l) ? vbuf[r - s] : 0.0;
      }
    }
    for (int j0 = 0; warp + QR_WARPS * j0 < ncols; j0 += 32) {
      double myrs = 0.0;      // lane j: R[s, column j of this chunk]
      bool myact = false;
      const int jend = min(32, (ncols - warp + QR_WARPS - 1) / QR_WARPS - j0);
      for (int jj = 0; jj < jend; jj += CB) {
        bool act[CB];
        int lcs[CB];
        bool all_sm = true;
#pragma unroll
        for (int cc = 0; cc < CB; ++cc) {
          lcs[cc] = warp + QR_WARPS * (j0 + jj + cc);
          act[cc] = (jj + cc < jend) && (lpos[lcs[cc]] > s);
          all_sm = all_sm && (!act[cc] || lcs[cc] < csm);
        }
        // SM = true: every live column of the batch sits in the shared-memory cache, and the pointers are
        // formed from `cache` alone so that ptxas emits LDS/STS instead of generic LD/ST.
        auto batch = [&](auto sm_tag) {
          constexpr bool SM = decltype(sm_tag)::value;
          double* a[CB];
#pragma unroll
          for (int cc = 0; cc < CB; ++cc) {
            if (SM) a[cc] = cache + (size_t)(act[cc] ? lcs[cc] : 0) * l;
            else a[cc] = act[cc] ? colptr(lcs[cc]) : vbuf;
          }
          if (NR > 0) {
            double ar[CB][NR > 0 ? NR : 1];
            double dot[CB];
#pragma unroll
            for (int cc = 0; cc < CB; ++cc) {
              double d0 = 0.0, d1 = 0.0;
#pragma unroll
              for (int i = 0; i < NR; ++i) {
                const int r = s + lane + 32 * i;
                ar[cc][i] = (act[cc] && r < l) ? a[cc][r] : 0.0;
                if (i & 1) d1 = fma(ar[cc][i], vr[i], d1);
                else d0 = fma(ar[cc][i], vr[i], d0);
              }
              dot[cc] = d0 + d1;
            }
            // packed butterflies: each sum keeps the association of the plain xor butterfly (bitwise the same result),
            // but CB sums cost CB+2 (CB = 4) or 6 (CB = 2) exchanges instead of 5 CB
            if (CB == 4) {
              const bool h16 = lane & 16, h8 = lane & 8;
              const double wa = (h16 ? dot[2] : dot[0]) + __shfl_xor_sync(0xffffffffu, h16 ? dot[0] : dot[2], 16);
              const double wb = (h16 ? dot[3] : dot[1]) + __shfl_xor_sync(0xffffffffu, h16 ? dot[1] : dot[3], 16);
              double w = (h8 ? wb : wa) + __shfl_xor_sync(0xffffffffu, h8 ? wa : wb, 8);
              w += __shfl_xor_sync(0xffffffffu, w, 4);
              w += __shfl_xor_sync(0xffffffffu, w, 2);
              w += __shfl_xor_sync(0xffffffffu, w, 1);
              // sum 0 in lanes 0-7, sum 1 in 8-15, sum 2 in 16-23, sum 3 in 24-31
#pragma unroll
              for (int cc = 0; cc < 4; ++cc) dot[cc] = __shfl_sync(0xffffffffu, w, 8 * cc);
            } else if (CB == 2) {
              const bool h16 = lane & 16;
              double w = (h16 ? dot[1] : dot[0]) + __shfl_xor_sync(0xffffffffu, h16 ? dot[0] : dot[1], 16);
              w += __shfl_xor_sync(0xffffffffu, w, 8);
              w += __shfl_xor_sync(0xffffffffu, w, 4);
              w += __shfl_xor_sync(0xffffffffu, w, 2);
              w += __shfl_xor_sync(0xffffffffu, w, 1);
              dot[0] = __shfl_sync(0xffffffffu, w, 0);
              dot[1 % CB] = __shfl_sync(0xffffffffu, w, 16);
            } else {
#pragma unroll
              for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
                for (int cc = 0; cc < CB; ++cc) dot[cc] += __shfl_xor_sync(0xffffffffu, dot[cc], o);
              }
            }
#pragma unroll
            for (int cc = 0; cc < CB; ++cc) {
              const double f = tau * dot[cc];
#pragma unroll
              for (int i = 0; i < NR; ++i) {
                const int r = s + lane + 32 * i;
                ar[cc][i] = fma(-f, vr[i], ar[cc][i]);
                if (act[cc] && r < l) a[cc][r] = ar[cc][i];
              }
              const double rs = __shfl_sync(0xffffffffu, ar[cc][0], 0);
              if (lane == jj + cc) {
                myrs = rs;
                myact = act[cc];
              }
            }
          } else {
#pragma unroll
            for (int cc = 0; cc < CB; ++cc) {
              if (!act[cc]) {
                if (lane == jj + cc) myact = false;
                continue;
              }
              double dot = 0.0;
              for (int r = s + lane; r < l; r += 32) dot = fma(a[cc][r], vbuf[r - s], dot);
              dot = warp_sum(dot);
              const double f = tau * dot;
              double rs = 0.0;
              for (int r = s + lane; r < l; r += 32) {
                const double x = fma(-f, vbuf[r - s], a[cc][r]);
                a[cc][r] = x;
                if (r == s) rs = x;
              }
              rs = __shfl_sync(0xffffffffu, rs, 0);
              if (lane == jj + cc) {
                myrs = rs;
                myact = true;
              }
            }
          }
        };
        if (all_sm) batch(std::true_type{});
        else batch(std::false_type{});
      }
      // LAWN-176 downdate, one column per lane (dlaqps step 8)
      bool flagged = false;
      const int mylc = warp + QR_WARPS * (j0 + lane);
      if (downdate && myact) {
        const double v1 = vn1[mylc];
        if (v1 != 0.0) {
          double t = fabs(myrs) / v1;
          t = fmax(0.0, (1.0 + t) * (1.0 - t));
          const double q = v1 / vn2[mylc];
          const double t2 = t * (q * q);
          if (t2 <= TOL3Z) flagged = true;
          else vn1[mylc] = v1 * sqrt(t);
        }
      }
      unsigned fm = __ballot_sync(0xffffffffu, flagged);
      if (fm) wflag = 1;
      while (fm) {
        // flagged: dlaqps ends the block and recomputes the norm from rows s+1..l-1 of the updated column
        const int fl = __ffs(fm) - 1;
        fm &= fm - 1;
        const int lc = warp + QR_WARPS * (j0 + fl);
        const double* a = colptr(lc);
        double ss2 = 0.0;
        for (int r = s + 1 + lane; r < l; r += 32) ss2 = fma(a[r], a[r], ss2);
        ss2 = warp_sum(ss2);
        if (lane == 0) {
          const double nn = sqrt(ss2);
          vn1[lc] = nn;
          vn2[lc] = nn;
        }
      }
    }
    __syncwarp();
    QR_TICK(4)

    // ---- end of step ----
    ++cnt;
    ++s;
    bool block_end = (cnt == jb);
    if (block_end) {
      // block ends by count; flags raised in this step are irrelevant
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
    }
    warp_candidate(s, block_end ? 0 : wflag);
    __syncthreads();
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
    for (int i = 0; i < 5; ++i) p.info[4 + i] = (int)(s_tph[i] >> 10);   // kilo-cycles per phase
  }
  if (tid == 0 && p.dbg)
    for (int i = 0; i < 5; ++i) p.dbg[cta * 8 + i] = (int)(s_tph[i] >> 10);
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
                                double* __restrict__ R12, int64_t ld12) {
  // R = triu(B[0:k, p]) split into R11 (k x k, ld k) and R12 (k x (n-k), ld k)  (src/pqr.jl:428)
  for (int64_t j = blockIdx.x; j < n; j += gridDim.x) {
    const double* s = B + (jpvt1[j] - 1) * ldb;
    if (j < k) {
      double* d = R11 + j * (int64_t)k;
      for (int r = threadIdx.x; r < k; r += blockDim.x) d[r] = (r <= j) ? s[r] : 0.0;
    } else {
      double* d = R12 + (j - k) * ld12;
      for (int r = threadIdx.x; r < k; r += blockDim.x) d[r] = s[r];
      if (threadIdx.x == 0 && ld12 > k) d[k] = 0.0;      // padding row
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
                 double rtol, QrcpOut* out, bool nopivot) {
  out->k = 0;
  out->nsteps = 0;
  out->nblocks = 0;
  out->status = 0;
  if (kcap <= 0 || n <= 0 || l <= 0) return BRA_OK;
  const int nbe = nb < kcap ? nb : kcap;
  // tall problems (more rows than the on-chip kernel of qrcp_fast.cu takes): the blocked dlaqps with the trailing update
  // on the FP64 tensor cores (qrcp_blocked.cu)
  if (!nopivot && l > 576 && bra_qrcp_blocked_ok(l, n, nbe, ctx->num_sms))
    return bra_qrcp_blocked_run(ctx, B, ldb, l, n, kcap, nb, atol, rtol, out);

  // grid: one CTA per SM, but keep >= 8 columns per CTA
  int G = ctx->num_sms < MAXG ? ctx->num_sms : MAXG;
  int64_t maxG = (n + 7) / 8;
  if (maxG < G) G = (int)(maxG < 1 ? 1 : maxG);
  const int cpc = (int)((n + G - 1) / G);

  // shared-memory budget
  const size_t fixed = ((size_t)l + nbe + MAXG + QR_WARPS) * 8 + 2 * MAXG * 4 + 2 * QR_WARPS * sizeof(Cand) + 64;
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
  // LL exchange buffers: [candidate columns | header inboxes]; zeroed when (re)allocated or when the
  // 32-bit stamp epoch is about to wrap, otherwise reused across launches with a fresh epoch.
  const size_t col_bytes = (((size_t)2 * G * (((l + 1) & ~1) + RECH) * sizeof(LL16)) + 31) & ~size_t(31);
  const size_t inbox_bytes = (size_t)2 * MAXG * MAXG * sizeof(LL32);   // the fast kernel pads the inbox rows to MAXG
  const size_t rec_bytes = col_bytes + inbox_bytes;
  if (ctx->rec.cap < rec_bytes || ctx->rec_zeroed < ctx->rec.cap || ctx->rec_epoch > 0xF0000000u) {
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
  p.nopivot = nopivot ? 1 : 0;
  p.atol = atol;
  p.rtol = rtol;
  p.cpc = cpc;
  p.csm = csm;
  p.meta_smem = meta_smem;
  p.vn1g = ctx->vn1.as<double>();
  p.vn2g = ctx->vn2.as<double>();
  p.lposg = ctx->lpos.as<int>();
  p.rec = ctx->rec.as<LL16>();
  p.inbox = reinterpret_cast<LL32*>(reinterpret_cast<unsigned char*>(ctx->rec.p) + col_bytes);
  p.epoch = ctx->rec_epoch;
  p.jpvt = ctx->jpvt.as<int64_t>();
  p.tau = ctx->tau.as<double>();
  p.rdiag = ctx->rdiag.as<double>();
  p.info = ctx->info.as<int>();
  p.kbtrace = ctx->kbtrace.as<int>();
  p.kbcap = kcap + 1;
  BRA_CUDA(ctx->scratch3.reserve((size_t)MAXG * 8 * 4));
  p.dbg = ctx->scratch3.as<int>();
  p.lds = l;
  p.fast = 0;
  p.ts = nullptr;
  p.ts_step = -1;
  if (const char* ev = getenv("BRA_QRCP_TS_STEP")) {
    BRA_CUDA(ctx->scratch3.reserve((size_t)MAXG * 8 * 4 + (size_t)MAXG * QR_WARPS * 16 * 8));
    p.dbg = ctx->scratch3.as<int>();
    p.ts = reinterpret_cast<long long*>(reinterpret_cast<unsigned char*>(ctx->scratch3.p) + (size_t)MAXG * 8 * 4);
    p.ts_step = atoi(ev);
    BRA_CUDA(cudaMemsetAsync(p.ts, 0, (size_t)MAXG * QR_WARPS * 16 * 8, ctx->stream));
  }
  ctx->rec_epoch += (uint32_t)l + 8;

  cudaError_t e;
  static const bool use_v1 = getenv("BRA_QRCP_V1") != nullptr;
  // short sketches (l <= 576, <= 120 columns per CTA): the warp-specialised kernel of qrcp_fast.cu
  const bool aligned = ((reinterpret_cast<uintptr_t>(B) & 15) == 0) && (ldb % 2 == 0);
  int jw = 0, lds_f = 0, csm_f = 0;
  size_t smem_f = 0;
  const bool use_fast = !use_v1 && bra_qrcp_fast_plan(l, cpc, nbe, budget, aligned, &jw, &lds_f, &csm_f, &smem_f);
  if (use_fast) {
    p.lds = lds_f;
    p.csm = csm_f;
    p.meta_smem = 1;
    p.fast = getenv("BRA_QRCP_NOTMEM") ? 1 : 2;      // 2: the slab lives in tensor memory first
    e = bra_qrcp_fast_launch(p, G, jw, smem_f, ctx->stream);
  } else if (l <= 64) e = launch_qrcp<2>(p, G, smem, ctx->stream);
  else if (l <= 96) e = launch_qrcp<3>(p, G, smem, ctx->stream);
  else if (l <= 160) e = launch_qrcp<5>(p, G, smem, ctx->stream);
  else if (l <= 288) e = launch_qrcp<9>(p, G, smem, ctx->stream);
  else if (l <= 544) e = launch_qrcp<17>(p, G, smem, ctx->stream);
  else e = launch_qrcp<0>(p, G, smem, ctx->stream);
  BRA_CUDA(e);
  ctx->launches++;
  BRA_CUDA(cudaMemcpyAsync(ctx->h_info, ctx->info.p, 48, cudaMemcpyDeviceToHost, ctx->stream));
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
                 double* R11, double* R12, int64_t ld12) {
  if (n <= 0 || k <= 0) return BRA_OK;
  int grid = (int)(n < 148 * 8 ? n : 148 * 8);
  gather_R_kernel<<<grid, 128, 0, ctx->stream>>>(B, ldb, n, k, jpvt1, R11, R12, ld12);
  ctx->launches++;
  BRA_CUDA(cudaGetLastError());
  return BRA_OK;
}

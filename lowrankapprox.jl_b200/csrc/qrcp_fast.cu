// Early-terminating Householder QRCP of a SHORT sketch (l <= 576 rows): the warp-specialised persistent kernel.
// Semantics: geqp3_adap_main! + LAPACK dlaqps (reference: src/pqr.jl:361-418, src/lapack.jl:117-139), exactly as
// the general kernel of qrcp.cu (first-maximum pivots, norm hand-over on a swap, LAWN-176 downdate, a flagged column
// ends the block, rank test at block ends only, pivoting continues to the block end).
//
// Why a second kernel: measured on B200 (tools/gpu_qrcp_trace.py, ncu source counters) the pivot step is bound by
// DEPENDENT-INSTRUCTION latency -- about 8 cycles per instruction per warp, ~1250 instructions per warp and step in the
// general kernel -- not by bandwidth.  So the critical chain of a step is cut to the bone and everything else runs
// beside it:
//   * 15 COMPUTE warps own the columns (lc = warp + 15 j, at most JW = 4 or 8 per warp: straight-line code, no loops).
//     Fixed row map: lane owns rows (64 c + 2 lane, +1) of every 64-row chunk c, so every slab access is a 128-bit
//     LDS/STS, v lives in 2 NCH registers per lane and the code is re-dispatched on the number of live chunks
//     NCH = ceil(l/64) - floor(s/64), which shrinks as the factorization proceeds.
//       pass 1 (critical, read-only): dots v.a_j -> f_j = tau v.a_j, the pivot-row entry R[s,j] = a_j[s] - f_j, the
//              LAWN-176 downdate in a division-free form (1/vn1 and (vn1/vn2)^2 are kept per column and refreshed
//              OFF the chain) and the candidate key vn1^2 temp, which orders like the downdated norm: no square root,
//              no division on the chain;
//       pass 2 (hidden behind the header exchange): a_j -= f_j v.  The CTA's candidate column goes first: its owner
//              warp alone finishes it, runs dlarfg for the next step and publishes the Householder vector.
//   * 1 COMM warp: merges the 15 warp candidates, pushes one 32-byte header into every CTA's inbox, polls its own
//     inbox, picks the winner, keeps the dlaqps block bookkeeping, fetches the winner's record into shared memory
//     (double buffered: it runs ahead of pass 2) and stores the winner column in LAPACK layout.
//   Two named barriers per step; shared arrays are addressed by 32-bit offsets (no generic pointers in registers).
#include "common.cuh"
#include <algorithm>
#include <cstdlib>
#include <type_traits>
#include "qrcp_common.cuh"
#include "qrcp_exchange.cuh"

namespace {

constexpr int QF_MAXCH = 9;               // 64-row chunks: l <= 576
constexpr int QF_THREADS = QR_THREADS;    // 512: 15 compute warps + the comm warp (registers are allocated per 4 warps: fewer
                                          // threads would not buy more registers until 384)
constexpr int QF_WARPS = QF_THREADS / 32;
constexpr int QF_CW = QF_WARPS - 1;       // compute warps
constexpr long long QF_NONE = -1;         // "no candidate" key (valid keys are bit patterns of non-negative doubles)

struct QfCtrl {
  double tau;
  int stop;       // after the fetch: 0: apply step s and go on; 1: apply step s in full, then exit
  int nsteps;
  int pre;        // before the fetch: 0: fetch the record of CTA wcta; 2: exit now (rank found); 3: exchange failure
  int wcta;
  int fail;       // a fetching warp timed out
  int pad;
};
struct __align__(16) QfSlotA {     // one LDS.128 / STS.128
  long long key;
  int lp, ps;
};
struct __align__(8) QfSlotB {
  int lc, flag;
};

__device__ __forceinline__ void qf_bar(int id) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(QF_THREADS) : "memory"); }

// argmax of (key, lp) over the warp (key: i64, larger wins; ties: smaller lp); returns the winning lane
__device__ __forceinline__ int warp_argmax_key(long long key, int lp) {
  const int hi = (int)(key >> 32);
  const unsigned lo = (unsigned)key;
  const int mh = __reduce_max_sync(0xffffffffu, hi);
  const bool c1 = (hi == mh);
  const unsigned ml = __reduce_max_sync(0xffffffffu, c1 ? lo : 0u);
  const bool c2 = c1 && (lo == ml);
  const int mlp = __reduce_min_sync(0xffffffffu, c2 ? lp : 0x7fffffff);
  return __ffs(__ballot_sync(0xffffffffu, c2 && lp == mlp)) - 1;
}

// ---- tensor memory as the home of the slab.  Measured on B200 (tools/tmem_probe.cu): tcgen05.ld 32x32b streams at
// ~810 B/clk/SM and tcgen05.st at ~750 B/clk/SM with 16 warps -- six times the 128 B/clk of shared memory -- on a pipe
// of its own (LDS and TMEM loads overlap fully), 23 cycles dependent-load latency, 256 KB per SM.  A warp reaches the 32
// TMEM lanes of its quadrant (warp % 4); thread t of the warp owns lane t, so "lane owns rows (64 c + 2 lane, +1)" maps
// one 64-row chunk of a column onto 4 consecutive 32-bit TMEM columns (x4 shape: two doubles per thread). ----
__device__ __forceinline__ void tm_ld2(uint32_t taddr, double2& x) {
  uint32_t r0, r1, r2, r3;
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(taddr)
               : "memory");
  // NOTE: valid only after tm_wait_ld(); the conversions below are register pairings, scheduled after the wait by the
  // volatile ordering of the two asm statements
  x.x = __hiloint2double((int)r1, (int)r0);
  x.y = __hiloint2double((int)r3, (int)r2);
}
__device__ __forceinline__ void tm_st2(uint32_t taddr, const double2 x) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};" ::"r"(taddr), "r"((uint32_t)__double2loint(x.x)),
               "r"((uint32_t)__double2hiint(x.x)), "r"((uint32_t)__double2loint(x.y)), "r"((uint32_t)__double2hiint(x.y))
               : "memory");
}
__device__ __forceinline__ void tm_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tm_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// Tensor-memory budget of a compute warp.  A warp reaches the 32 lanes of quadrant warp % 4 over all 512 columns; the
// quadrant is shared by the compute warps w, w + 4, w + 8, (w + 12) -- four in quadrants 0-2, three in quadrant 3 (the comm
// warp keeps no columns).  The quadrant's floor(512 / tchunk) slab columns are dealt out as evenly as possible (round 2a:
// fixed 128- / 168-column windows wasted up to 20 TMEM columns per warp: 48 instead of 56 slab columns at l = 520).
__host__ __device__ __forceinline__ int qf_tcap(int warp, int tchunk) {        // slab columns of this warp in TMEM
  const int nw = (warp & 3) < 3 ? 4 : 3, r = warp >> 2, cq = 512 / tchunk;
  return cq / nw + (r < cq % nw ? 1 : 0);
}
__host__ __device__ __forceinline__ int qf_toff(int warp, int tchunk) {        // first TMEM column of this warp's share
  const int nw = (warp & 3) < 3 ? 4 : 3, r = warp >> 2, cq = 512 / tchunk;
  return tchunk * (r * (cq / nw) + (r < cq % nw ? r : cq % nw));
}

#ifdef BRA_QRCP_TRACE
__device__ __forceinline__ long long qf_gtime() {
  long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define QF_TS(k)                                                                                   \
  if (p.ts && lane == 0 && s == p.ts_step) {                                                       \
    p.ts[((size_t)cta * QR_WARPS + warp) * 16 + (k)] = clock64();                                  \
    if ((k) < 4) p.ts[((size_t)cta * QR_WARPS + warp) * 16 + 9 + (k)] = qf_gtime();                \
  }
#define QF_TICK(i) { long long _t = clock64(); s_tph[i] += _t - tlast; tlast = _t; }
#else
#define QF_TS(k)
#define QF_TICK(i)
#endif

// NMAX = ceil(l / 64) bound of the launch: instances above it are never generated, so the register peak of the long
// instances (v in 2 NCH registers per lane) does not spill the step loop of the short, latency-critical sketches
#define QF_DISPATCH(nchv, FN, ...)                                                                     \
  switch (nchv) {                                                                                      \
    case 1: FN(std::integral_constant<int, 1>{}, __VA_ARGS__); break;                                  \
    case 2: if constexpr (NMAX >= 2) FN(std::integral_constant<int, 2>{}, __VA_ARGS__); break;         \
    case 3: if constexpr (NMAX >= 3) FN(std::integral_constant<int, 3>{}, __VA_ARGS__); break;         \
    case 4: if constexpr (NMAX >= 4) FN(std::integral_constant<int, 4>{}, __VA_ARGS__); break;         \
    case 5: if constexpr (NMAX >= 5) FN(std::integral_constant<int, 5>{}, __VA_ARGS__); break;         \
    case 6: if constexpr (NMAX >= 6) FN(std::integral_constant<int, 6>{}, __VA_ARGS__); break;         \
    case 7: if constexpr (NMAX >= 7) FN(std::integral_constant<int, 7>{}, __VA_ARGS__); break;         \
    case 8: if constexpr (NMAX >= 8) FN(std::integral_constant<int, 8>{}, __VA_ARGS__); break;         \
    default: if constexpr (NMAX >= 9) FN(std::integral_constant<int, 9>{}, __VA_ARGS__); break;        \
  }

template <int JW, int NMAX>
__global__ void __launch_bounds__(QF_THREADS, 1) qrcp_fast_kernel(QrcpParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int G = gridDim.x, cta = blockIdx.x;
  const int l = p.l;
  const int lds = p.lds;
  const int64_t col0 = (int64_t)cta * p.cpc;
  const int ncols = (int)max((int64_t)0, min((int64_t)p.cpc, p.n - col0));
  const int csm = min(p.csm, ncols);
  const int nchtot = (l + 63) >> 6;
  const int LV = nchtot << 6;
  const int cpe = (p.cpc + 1) & ~1;

  // ---- shared memory carve-up (every array starts 16-byte aligned; 32-bit byte offsets from the dynamic base) ----
  const int off_ctrl = 2 * LV * 8;                    // vbuf [2][LV]: Householder vector by ABSOLUTE row, zero padded;
                                                      //   step s uses buffer s & 1 (the comm warp runs ahead of pass 2)
  const int off_rd = off_ctrl + 32;                   // rdblk: nb (even) diagonal entries of the current block
  const int off_sa = off_rd + ((p.nb + 1) & ~1) * 8;  // warp candidates: QfSlotA[16], QfSlotB[16]
  const int off_sb = off_sa + 16 * 16;
  const int off_lv = off_sb + 16 * 8;                 // slive[16]: live-column mask of each compute warp for the current step
  const int off_f = off_lv + 16 * 4;                  // fbuf: f_j of the current step                        cpe doubles
  const int off_st = off_f + cpe * 8;                 // st2: {1/vn1, (vn1/vn2)^2} per column                 cpe double2
  const int off_sq = off_st + cpe * 16;               // ssq = vn1^2, sv1 = vn1, srv2 = 1/vn2, stmp = downdate factor of
                                                      //   the current step (< 0: nothing to refresh)            4 cpe doubles
  const int off_lpos = off_sq + 4 * cpe * 8;          // lpos: logical LAPACK position per column             cpe ints (x4)
  const int off_cache = off_lpos + ((cpe + 3) & ~3) * 4;   // the slab: csm columns of lds doubles
#define vbuf reinterpret_cast<double*>(smem_raw)
#define ctrl reinterpret_cast<QfCtrl*>(smem_raw + off_ctrl)
#define rdblk reinterpret_cast<double*>(smem_raw + off_rd)
#define slotA reinterpret_cast<QfSlotA*>(smem_raw + off_sa)
#define slotB reinterpret_cast<QfSlotB*>(smem_raw + off_sb)
#define slive reinterpret_cast<unsigned*>(smem_raw + off_lv)
#define fbuf reinterpret_cast<double*>(smem_raw + off_f)
#define st2 reinterpret_cast<double2*>(smem_raw + off_st)
#define ssq reinterpret_cast<double*>(smem_raw + off_sq)
#define sv1 (ssq + cpe)
#define srv2 (ssq + 2 * cpe)
#define stmp (ssq + 3 * cpe)
#define lpos reinterpret_cast<int*>(smem_raw + off_lpos)
#define cache reinterpret_cast<double*>(smem_raw + off_cache)
  const int recs = lds + RECH;                        // record stride (even: 32-byte aligned row pairs)
#ifdef BRA_QRCP_TRACE
  __shared__ long long s_tph[6];                      // per-phase cycle totals (diagnostic)
  if (tid < 6) s_tph[tid] = 0;
#endif

  for (int r = tid; r < 2 * LV; r += QF_THREADS) vbuf[r] = 0.0;
  if (tid == 0) {
    ctrl->tau = 0.0;
    ctrl->stop = 0;
    ctrl->nsteps = 0;
    ctrl->pre = 0;
    ctrl->wcta = 0;
    ctrl->fail = 0;
  }

  // ---- tensor memory: all 512 columns, allocated (and released at the end) by the comm warp ----
  __shared__ uint32_t s_tbase;
  if (warp == QF_CW) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(
                     (uint32_t)__cvta_generic_to_shared(&s_tbase))
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tbase = s_tbase;

  // ---- where column j of compute warp w lives: class 0 = tensor memory (the first tcap columns of every warp),
  //      class 1 = shared memory slot (j - tcap) * 15 + w while slots last, class 2 = global memory (L2) ----
  const int tchunk = 4 * nchtot;                       // TMEM columns per slab column
  auto tcap_of = [&](int w) { return p.fast >= 2 ? qf_tcap(w, tchunk) : 0; };
  auto tm_col = [&](int w, int j) -> uint32_t {        // TMEM address of chunk 0 of column j of warp w
    return tbase + ((uint32_t)(32 * (w & 3)) << 16) + (uint32_t)(qf_toff(w, tchunk) + j * tchunk);
  };
  auto class_of = [&](int w, int j, int& slot) -> int {
    const int tc = tcap_of(w);
    if (j < tc) {
      slot = j;
      return 0;
    }
    slot = (j - tc) * QF_CW + w;
    return slot < p.csm ? 1 : 2;
  };

  // ---- prologue: every compute warp stages its own columns (only the owner reaches its TMEM lanes) and takes the
  //      initial column norms (src/pqr.jl:376-385) ----
  if (warp < QF_CW) {
    for (int j = 0; warp + QF_CW * j < ncols; ++j) {
      const int lc = warp + QF_CW * j;
      const double* g = p.B + (col0 + lc) * p.ldb;
      int slot;
      const int cls = class_of(warp, j, slot);
      double amax = 0.0;
      for (int r = lane; r < l; r += 32) amax = fmax(amax, fabs(g[r]));
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) amax = fmax(amax, __shfl_xor_sync(0xffffffffu, amax, o));
      // exact power-of-two scaling: same rounding as the unscaled sum, no overflow/underflow
      const int e = amax > 0.0 ? ilogb(amax) : 0;
      const double sc = scalbn(1.0, -e);
      double ss = 0.0;
      for (int c = 0; c < nchtot; ++c) {
        const int r = 64 * c + 2 * lane;
        double2 x;
        x.x = r < l ? g[r] : 0.0;
        x.y = r + 1 < l ? g[r + 1] : 0.0;
        ss = fma(x.x * sc, x.x * sc, ss);
        ss = fma(x.y * sc, x.y * sc, ss);
        if (cls == 0) tm_st2(tm_col(warp, j) + 4 * c, x);
        else if (cls == 1 && r < lds) *reinterpret_cast<double2*>(cache + (size_t)slot * lds + r) = x;
      }
      ss = warp_sum(ss);
      const double nrm = amax > 0.0 ? scalbn(sqrt(ss), e) : 0.0;
      if (lane == 0) {
        const double rn = nrm != 0.0 ? 1.0 / nrm : 0.0;
        st2[lc] = make_double2(rn, 1.0);
        ssq[lc] = nrm * nrm;
        sv1[lc] = nrm;
        srv2[lc] = rn;
        lpos[lc] = (int)(col0 + lc);
      }
    }
    tm_wait_st();
  }
  __syncthreads();

  // ---- fetch of one 64-row chunk of the winner's record (v already scaled by its owner) into vbuf, by absolute row:
  //      0 below s, 1 at s, v above.  In the winner CTA the same warp stores the chunk of the pivot column in LAPACK
  //      layout (beta at row s, v below).  Every warp takes one chunk: the record is fetched in ONE L2 round trip. ----
  auto fetch_chunk = [&](const int s, const int wcta, const int c) {
    const int par = s & 1;
    const uint32_t stamp = p.epoch + (uint32_t)s;
    const LL16* wrec = p.rec + ((size_t)par * G + wcta) * recs;
    const int r = ((s >> 6) << 6) + 64 * c + 2 * lane;
    const bool need = r + 1 > s && r < l;
    const bool isw = wcta == cta;
    uint32_t q[8], hb[4], hp[4], hs[4];
    if (need) ll32_ld(reinterpret_cast<const LL32*>(wrec + RECH + r), q);
    ll_ld(wrec + 3, hs[0], hs[1], hs[2], hs[3]);
    if (isw) {
      ll_ld(wrec + 1, hb[0], hb[1], hb[2], hb[3]);
      ll_ld(wrec + 2, hp[0], hp[1], hp[2], hp[3]);
    }
    uint32_t spins = 0;
    while (true) {
      bool ok = true;
      if (need) {
        const bool ok0 = (r <= s) || ((q[1] ^ stamp) | (q[3] ^ stamp)) == 0;
        const bool ok1 = (r + 1 >= l) || ((q[5] ^ stamp) | (q[7] ^ stamp)) == 0;
        if (!(ok0 && ok1)) {
          ok = false;
          ll32_ld(reinterpret_cast<const LL32*>(wrec + RECH + r), q);
        }
      }
      if (((hs[1] ^ stamp) | (hs[3] ^ stamp)) != 0) {
        ok = false;
        ll_ld(wrec + 3, hs[0], hs[1], hs[2], hs[3]);
      }
      if (isw && (((hb[1] ^ stamp) | (hb[3] ^ stamp) | (hp[1] ^ stamp) | (hp[3] ^ stamp)) != 0)) {
        ok = false;
        ll_ld(wrec + 1, hb[0], hb[1], hb[2], hb[3]);
        ll_ld(wrec + 2, hp[0], hp[1], hp[2], hp[3]);
      }
      if (__all_sync(0xffffffffu, ok)) break;
      if (++spins > SPIN_LIMIT) {
        if (lane == 0) ctrl->fail = 1;
        return;
      }
    }
    if (r < l) {
      const double scale = __hiloint2double((int)hs[2], (int)hs[0]);
      double2 v;
      v.x = r > s ? __hiloint2double((int)q[2], (int)q[0]) * scale : (r == s ? 1.0 : 0.0);
      v.y = r + 1 > s ? __hiloint2double((int)q[6], (int)q[4]) * scale : (r + 1 == s ? 1.0 : 0.0);
      if (r + 1 >= l) v.y = 0.0;
      *reinterpret_cast<double2*>(vbuf + par * LV + r) = v;
      if (isw) {
        // the pivot column in LAPACK layout: into its shared-memory slot, or straight to its final place in global
        // memory when it lives in tensor memory (only its owner warp could write it there) or in L2
        const double beta = __hiloint2double((int)hb[2], (int)hb[0]);
        const int wlc = (int)((int)hp[0] - col0);
        int wslot;
        const bool wsm = class_of(wlc % QF_CW, wlc / QF_CW, wslot) == 1;
        double* wa = wsm ? cache + (size_t)wslot * lds : p.B + (col0 + wlc) * p.ldb;
        if (r >= s) {
          const double x0 = r == s ? beta : v.x;
          if (wsm) wa[r] = x0;
          else __stcg(wa + r, x0);
        }
        if (r + 1 >= s && r + 1 < l) {
          const double x1 = r + 1 == s ? beta : v.y;
          if (wsm) wa[r + 1] = x1;
          else __stcg(wa + r + 1, x1);
        }
      }
    }
  };

  const int lastrk = (int)min((int64_t)l, p.n);
  int s = 0;                                  // current pivot step (every warp keeps its own copy)

  if (warp == QF_CW) {
    // =================================== COMM warp ===================================
    int jblk = 0, cnt = 0, jb = min(p.nb, p.kcap), nblocks = 0, kres = -1;
    double ptol = 0.0;
    bool failed = false, prev_block_end = true;     // no flags exist before step 0
#ifdef BRA_QRCP_TRACE
    long long tlast = clock64();
#endif
    // the inbox rows are padded to MAXG words, so the five words of a lane (CTAs lane + 32 i) sit at compile-time
    // offsets from one base address; lanes whose CTA does not exist alias a valid slot and ignore it
    bool have[5];
#pragma unroll
    for (int i = 0; i < 5; ++i) have[i] = lane + 32 * i < G;
    constexpr size_t PUSH_STRIDE = (size_t)32 * MAXG;      // LL32 words between destinations lane + 32 i and lane + 32 (i+1)
    constexpr size_t PAR_STRIDE = (size_t)MAXG * MAXG;

    qf_bar(2);                                      // candidates for step 0
    while (true) {
      QF_TS(0)
      // ---- merge the warp candidates -> this CTA's candidate for step s ----
      long long key = QF_NONE;
      int lp = 0x7fffffff, psx = -1, fl = 0;
      if (lane < QF_CW) {
        const QfSlotA a = slotA[lane];
        key = a.key;
        lp = a.lp;
        psx = a.ps;
        fl = slotB[lane].flag;
      }
      const int wl = warp_argmax_key(key, lp);
      const long long ckey = __shfl_sync(0xffffffffu, key, wl);
      const int clp = __shfl_sync(0xffffffffu, lp, wl);
      const int my_ps = __reduce_max_sync(0xffffffffu, psx);     // physical column at logical position s, if owned here
      const int cflag = prev_block_end ? 0 : (__any_sync(0xffffffffu, fl != 0) ? 1 : 0);

      // ---- publish: one 32-byte header word into every CTA's inbox ----
      const uint32_t stamp = p.epoch + (uint32_t)s;
      const int par = s & 1;
      {
        LL32* dst = p.inbox + par * PAR_STRIDE + (size_t)lane * MAXG + cta;
#pragma unroll
        for (int i = 0; i < 5; ++i)
          if (have[i]) ll32_store(dst + i * PUSH_STRIDE, __longlong_as_double(ckey), clp, cflag, stamp);
      }
      QF_TS(1)
#ifdef BRA_QRCP_TRACE
      if (lane == 0) QF_TICK(0)
#endif

      // ---- gather my inbox: light polling loop (loads + stamp test only), the winner is picked once all are in ----
      uint32_t q[5][8];
      {
        const LL32* src = p.inbox + par * PAR_STRIDE + (size_t)cta * MAXG + min(lane, G - 1);   // lanes >= G alias slot G-1
        uint32_t spins = 0;
        while (true) {
          uint32_t bad = 0;
#pragma unroll
          for (int i = 0; i < 5; ++i) ll32_ld(src + (have[i] ? 32 * i : 0), q[i]);
#pragma unroll
          for (int i = 0; i < 5; ++i) bad |= (q[i][1] ^ stamp) | (q[i][3] ^ stamp) | (q[i][5] ^ stamp) | (q[i][7] ^ stamp);
          if (!__any_sync(0xffffffffu, bad != 0)) break;
          if (++spins > SPIN_LIMIT) {
            failed = true;
            break;
          }
        }
      }
      long long bkey = QF_NONE;
      int blp = 0x7fffffff, bsrc = -1;
      uint32_t bfl = 0;
#pragma unroll
      for (int i = 0; i < 5; ++i) {
        if (have[i]) {
          const long long k2 = ((long long)q[i][2] << 32) | (long long)q[i][0];
          const int lp2 = (int)q[i][4];
          bfl |= q[i][6];
          if (k2 > bkey || (k2 == bkey && lp2 < blp)) {
            bkey = k2;
            blp = lp2;
            bsrc = lane + 32 * i;
          }
        }
      }
      const int wl2 = warp_argmax_key(bkey, blp);
      const int wcta = __shfl_sync(0xffffffffu, bsrc, wl2);
      const long long gkey = __shfl_sync(0xffffffffu, bkey, wl2);
      const int lw = __shfl_sync(0xffffffffu, blp, wl2);          // winner's logical position
      const bool gflag = __any_sync(0xffffffffu, bfl != 0);
      failed = failed || wcta < 0;
      QF_TS(2)
#ifdef BRA_QRCP_TRACE
      if (p.ts && lane == 0 && s == p.ts_step) p.ts[((size_t)cta * QR_WARPS + warp) * 16 + 13] = wcta;
#endif
#ifdef BRA_QRCP_TRACE
      if (lane == 0) QF_TICK(2)
#endif
      int pre = failed ? 3 : 0;

      // ---- block bookkeeping for the previous step (needs the gathered flags) ----
      if (!pre && cnt > 0 && gflag) {
        // a column was flagged during step s-1: dlaqps ended its block there
        if (cta == 0 && lane == 0 && nblocks < p.kbcap) p.kbtrace[nblocks] = cnt;
        ++nblocks;
        const int jn = jblk + cnt;
        if (fabs(rdblk[cnt - 1]) <= ptol) {
          for (int i = 0; i < cnt; ++i)
            if (fabs(rdblk[i]) <= ptol) {
              kres = jblk + i;
              break;
            }
        }
        if (kres >= 0) pre = 2;
        jblk = jn;
        cnt = 0;
        jb = min(p.nb, p.kcap - jblk);
      }
      if (lane == 0) {
        ctrl->pre = pre;
        ctrl->wcta = wcta;
        if (pre) ctrl->nsteps = s;
      }
      qf_bar(3);                  // winner known: every warp fetches one chunk of its record
      if (pre) break;
      if (s == 0) ptol = fmax(p.atol, p.rtol * __longlong_as_double(gkey));   // src/pqr.jl:386-389 (step-0 keys are norms)

      // ---- tau, beta, physical column of the winner ----
      double tau, beta;
      int pw;
      {
        const LL16* wrec = p.rec + ((size_t)par * G + wcta) * recs;
        uint32_t h[3][4];
#pragma unroll
        for (int i = 0; i < 3; ++i) ll_ld(wrec + i, h[i][0], h[i][1], h[i][2], h[i][3]);
        uint32_t spins = 0;
        while (true) {
          bool ok = true;
#pragma unroll
          for (int i = 0; i < 3; ++i)
            if (((h[i][1] ^ stamp) | (h[i][3] ^ stamp)) != 0) {
              ok = false;
              ll_ld(wrec + i, h[i][0], h[i][1], h[i][2], h[i][3]);
            }
          if (ok) break;
          if (++spins > SPIN_LIMIT) {
            failed = true;
            break;
          }
        }
        tau = __hiloint2double((int)h[0][2], (int)h[0][0]);
        beta = __hiloint2double((int)h[1][2], (int)h[1][0]);
        pw = (int)h[2][0];
      }
      if (failed && lane == 0) ctrl->fail = 1;
      if (lane == 0 && !failed) {
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
      __syncwarp();
      // ---- end-of-step bookkeeping, known as soon as beta is (the rank test only looks at the diagonal) ----
      ++cnt;
      const bool block_end = (cnt == jb);
      if (block_end && !failed) {
        // block ends by count; flags raised in this step are irrelevant
        if (cta == 0 && lane == 0 && nblocks < p.kbcap) p.kbtrace[nblocks] = cnt;
        ++nblocks;
        const int jn = jblk + cnt;
        if (fabs(rdblk[cnt - 1]) <= ptol) {
          for (int i = 0; i < cnt; ++i)
            if (fabs(rdblk[i]) <= ptol) {
              kres = jblk + i;
              break;
            }
        }
        if (kres < 0) {
          jblk = jn;
          cnt = 0;
          if (jblk >= p.kcap) kres = p.kcap;
          else jb = min(p.nb, p.kcap - jblk);
        }
      }
      prev_block_end = block_end;
      const int stop = kres >= 0 ? 1 : 0;
      if (lane == 0) {
        ctrl->tau = tau;
        ctrl->stop = stop;
        ctrl->nsteps = s + 1;
      }
      QF_TS(3)
#ifdef BRA_QRCP_TRACE
      if (lane == 0) QF_TICK(3)
#endif
      qf_bar(1);                  // vbuf, tau, lpos are ready: the compute warps run pass 1 of step s
      if (ctrl->fail) {
        failed = true;
        break;
      }
      if (stop) break;
      ++s;
      qf_bar(2);                  // candidates for step s
    }
    if (cta == 0 && lane == 0) {
      p.info[0] = failed ? -1 : kres;
      p.info[2] = nblocks;
      p.info[3] = failed ? 1 : 0;
    }
  } else {
    // ================================= COMPUTE warps =================================
    const int jtot = ncols > warp ? min(JW, (ncols - warp + QF_CW - 1) / QF_CW) : 0;   // this warp's columns lc = warp + 15 j
    const bool mycol = lane < jtot;                                                     // lane j < jtot <-> column j
    const int mylc = mycol ? warp + QF_CW * lane : 0;
    const int tcap = tcap_of(warp);                                                     // columns j < tcap: tensor memory
    const uint32_t tcol0 = tm_col(warp, 0);
#ifdef BRA_QRCP_TRACE
    long long tlast = clock64();
#endif
    // residency of my column j (warp-uniform): 0 TMEM, 1 shared memory, 2 L2
    auto cls_of = [&](int j) { return j < tcap ? 0 : ((j - tcap) * QF_CW + warp < p.csm ? 1 : 2); };
    auto sm_col = [&](int j) -> double* { return cache + (size_t)((j - tcap) * QF_CW + warp) * lds; };
    auto gl_col = [&](int j) -> double* { return p.B + (col0 + warp + QF_CW * j) * p.ldb; };
    // chunks i0 .. i0+NCH-1 of my column j: lane gets rows (64 (i0 + c) + 2 lane, +1).  Rows >= l read as zero.
    auto load_col = [&](auto tag, const int j, const int cls, const int i0, double2* x) {
      constexpr int NCH = decltype(tag)::value;
      const int rbase = (i0 << 6) + 2 * lane;
      if (cls == 0) {
        const uint32_t ta = tcol0 + (uint32_t)(j * tchunk + 4 * i0);
#pragma unroll
        for (int c = 0; c < NCH; ++c) tm_ld2(ta + 4 * c, x[c]);
        tm_wait_ld();
      } else if (cls == 1) {
        const double2* a2 = reinterpret_cast<const double2*>(sm_col(j) + rbase);
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
          x[c] = make_double2(0.0, 0.0);
          if (c < NCH - 1 || rbase + 64 * c < l) x[c] = a2[32 * c];
        }
      } else {
        const double2* a2 = reinterpret_cast<const double2*>(gl_col(j) + rbase);
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
          x[c] = make_double2(0.0, 0.0);
          if (c < NCH - 1 || rbase + 64 * c < l) x[c] = __ldcg(a2 + 32 * c);
        }
      }
    };
    auto store_col = [&](auto tag, const int j, const int cls, const int i0, const double2* x) {
      constexpr int NCH = decltype(tag)::value;
      const int rbase = (i0 << 6) + 2 * lane;
      if (cls == 0) {
        const uint32_t ta = tcol0 + (uint32_t)(j * tchunk + 4 * i0);
#pragma unroll
        for (int c = 0; c < NCH; ++c) tm_st2(ta + 4 * c, x[c]);
      } else if (cls == 1) {
        double2* a2 = reinterpret_cast<double2*>(sm_col(j) + rbase);
#pragma unroll
        for (int c = 0; c < NCH; ++c)
          if (c < NCH - 1 || rbase + 64 * c < l) a2[32 * c] = x[c];
      } else {
        double2* a2 = reinterpret_cast<double2*>(gl_col(j) + rbase);
#pragma unroll
        for (int c = 0; c < NCH; ++c)
          if (c < NCH - 1 || rbase + 64 * c < l) __stcg(a2 + 32 * c, x[c]);
      }
    };

    // warp candidate -> slot
    auto publish = [&](long long key, int lp, int ps, int flag) {
      const int wl = warp_argmax_key(key, lp);
      const int bps = __reduce_max_sync(0xffffffffu, ps);
      if (lane == wl) {
        QfSlotA a;
        a.key = key;
        a.lp = lp;
        a.ps = bps;
        slotA[warp] = a;
        QfSlotB b;
        b.lc = key >= 0 ? mylc : -1;
        b.flag = flag;
        slotB[warp] = b;
      }
    };

    // ---- pass 1 of step s: dots, f_j, pivot-row entry, norm downdate, candidate for step s + 1 ----
    auto pass1 = [&](auto tag, const int s, const double tau, const bool downdate, const bool want_cand) {
      constexpr int NCH = decltype(tag)::value;
      const int par = s & 1;
      const int i0 = s >> 6;
      const int rbase = (i0 << 6) + 2 * lane;
      double2 vr[NCH];
      {
        const double2* v2 = reinterpret_cast<const double2*>(vbuf + par * LV + rbase);
#pragma unroll
        for (int c = 0; c < NCH; ++c) vr[c] = v2[32 * c];
      }
      // lane j: state of column j
      const int qpos = mycol ? lpos[mylc] : -1;
      const double2 stj = st2[mylc];
      const double sq = ssq[mylc];
      const unsigned live = __ballot_sync(0xffffffffu, qpos > s);
      if (lane == 0) slive[warp] = live;
      const int ls = (s & 63) >> 1;           // the lane that holds row s (first live chunk), component s & 1

      // dots, four columns per packed butterfly; the slowest residency class (the highest j) goes first
      double fmine = 0.0, myas = 0.0;
#pragma unroll
      for (int b = JW / 4 - 1; b >= 0; --b) {
        double dot[4];
#pragma unroll
        for (int cc = 3; cc >= 0; --cc) {
          const int j = 4 * b + cc;
          double d0 = 0.0, d1 = 0.0;
          if (j < jtot) {                         // warp-uniform
            double2 x[NCH];
            load_col(tag, j, cls_of(j), i0, x);
#pragma unroll
            for (int c = 0; c < NCH; ++c) {
              d0 = fma(x[c].x, vr[c].x, d0);
              d1 = fma(x[c].y, vr[c].y, d1);
            }
            // a_j[s] sits in lane ls of the first chunk; lane j keeps it
            const double asj = __shfl_sync(0xffffffffu, (s & 1) ? x[0].y : x[0].x, ls);
            if (lane == j) myas = asj;
          }
          dot[cc] = d0 + d1;
        }
        // packed butterfly: four sums for 6 exchanges; each keeps the association of the plain xor butterfly
        const bool h16 = lane & 16, h8 = lane & 8;
        const double wa = (h16 ? dot[2] : dot[0]) + __shfl_xor_sync(0xffffffffu, h16 ? dot[0] : dot[2], 16);
        const double wb = (h16 ? dot[3] : dot[1]) + __shfl_xor_sync(0xffffffffu, h16 ? dot[1] : dot[3], 16);
        double w = (h8 ? wb : wa) + __shfl_xor_sync(0xffffffffu, h8 ? wa : wb, 8);
        w += __shfl_xor_sync(0xffffffffu, w, 4);
        w += __shfl_xor_sync(0xffffffffu, w, 2);
        w += __shfl_xor_sync(0xffffffffu, w, 1);
        // the sum of column 4b + q sits in lanes 8q..8q+7: lane 8q stores f, lane j = 4b + q picks it up
        const double f = tau * w;
        const int jq = 4 * b + (lane >> 3);
        if ((lane & 7) == 0 && ((live >> jq) & 1)) fbuf[warp + QF_CW * jq] = f;
        const double fm = __shfl_sync(0xffffffffu, f, (lane & 3) << 3);
        if ((lane >> 2) == b) fmine = fm;
      }

      // LAWN-176 downdate (dlaqps step 8), one column per lane, division-free on the chain:
      //   t = |r| / vn1, temp = max(0, (1+t)(1-t)), temp2 = temp (vn1/vn2)^2; flagged iff temp2 <= tol3z,
      //   else vn1 *= sqrt(temp).  The candidate key vn1^2 temp orders like the downdated norm.
      long long key = QF_NONE;
      int lp = 0x7fffffff, ps = -1;
      bool flagged = false;
      double temp = -1.0;                 // >= 0: plain downdate, the norm state of my column is refreshed after bar 2
      if (qpos > s) {
        lp = qpos;
        if (qpos == s + 1) ps = (int)(col0 + mylc);
        if (p.nopivot) {
          key = qpos == s + 1 ? __double_as_longlong(1.0) : QF_NONE;
        } else if (downdate && sq != 0.0) {
          const double t = fabs(myas - fmine) * stj.x;
          const double tt = fmax(0.0, (1.0 + t) * (1.0 - t));
          flagged = tt * stj.y <= TOL3Z;
          if (!flagged) temp = tt;
          key = __double_as_longlong(sq * tt);
        } else {
          key = __double_as_longlong(sq);
        }
      }
      unsigned fmk = __ballot_sync(0xffffffffu, flagged);
      const int wflag = fmk != 0;
      while (fmk) {
        // flagged: dlaqps ends the block and recomputes the norm from rows s+1..l-1 of the UPDATED column, so this
        // column gets its pass 2 right here (f is cleared: the regular pass 2 skips it)
        const int fl = __ffs(fmk) - 1;
        fmk &= fmk - 1;
        const int lc = warp + QF_CW * fl;
        __syncwarp();
        const double f = fbuf[lc];
        const int cls = cls_of(fl);
        double2 x[NCH];
        load_col(tag, fl, cls, i0, x);
        double ss2 = 0.0;
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
          const int r = rbase + 64 * c;
          x[c].x = fma(-f, vr[c].x, x[c].x);
          x[c].y = fma(-f, vr[c].y, x[c].y);
          if (r > s) ss2 = fma(x[c].x, x[c].x, ss2);
          if (r + 1 > s) ss2 = fma(x[c].y, x[c].y, ss2);
        }
        store_col(tag, fl, cls, i0, x);
        ss2 = warp_sum(ss2);
        __syncwarp();
        const double nn = sqrt(ss2);
        if (lane == fl) {
          const double rn = nn != 0.0 ? 1.0 / nn : 0.0;
          key = __double_as_longlong(nn * nn);
          temp = -1.0;
          st2[lc] = make_double2(rn, 1.0);
          ssq[lc] = nn * nn;
          sv1[lc] = nn;
          srv2[lc] = rn;
          fbuf[lc] = 0.0;
        }
      }
      if (wflag) tm_wait_st();
      if (mycol) stmp[mylc] = temp;
      if (want_cand) publish(key, lp, ps, wflag);
    };

    // refresh of lane j's norm state after a plain downdate (OFF the candidate chain)
    auto refresh = [&](const double temp) {
      const double v1n = sv1[mylc] * sqrt(temp);
      const double qn = v1n * srv2[mylc];
      sv1[mylc] = v1n;
      ssq[mylc] = v1n * v1n;
      st2[mylc] = make_double2(v1n != 0.0 ? 1.0 / v1n : 0.0, qn * qn);
    };

    // ---- candidate column: finish step sp on it (if sp >= 0), then dlarfg for step sp+1 -> this CTA's record.
    //      ONE sweep: the updated entries go into the record unscaled while the sum of squares accumulates; tau, beta
    //      and the scaling 1/(alpha - beta) follow in the header (the receivers scale while they build v). ----
    auto cand_dlarfg = [&](auto tag, const int sp, const int cand_lc) {
      constexpr int NCH = decltype(tag)::value;
      const int sn = sp + 1;
      const uint32_t stamp = p.epoch + (uint32_t)sn;
      LL16* myrec = p.rec + ((size_t)(sn & 1) * G + cta) * recs + RECH;
      const int cj = cand_lc / QF_CW;           // my column index of the candidate
      const int cls = cls_of(cj);
      double f = 0.0;
      if (sp >= 0) {
        f = fbuf[cand_lc];
        __syncwarp();
        if (lane == 0) fbuf[cand_lc] = 0.0;
      }
      const int i0 = (sp >= 0 ? sp : 0) >> 6;
      const int rb = i0 << 6;
      const int rbase = rb + 2 * lane;
      double ss = 0.0, al = 0.0;
      double2 x[NCH];
      load_col(tag, cj, cls, i0, x);
      if (f != 0.0) {
        const double2* v2 = reinterpret_cast<const double2*>(vbuf + (sp & 1) * LV + rbase);
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
          const double2 v = v2[32 * c];
          x[c].x = fma(-f, v.x, x[c].x);
          x[c].y = fma(-f, v.y, x[c].y);
        }
        store_col(tag, cj, cls, i0, x);
      }
#pragma unroll
      for (int c = 0; c < NCH; ++c) {
        const int r = rbase + 64 * c;
        if (c < NCH - 1 || r < l) {
          // rows (r, r+1) are two adjacent LL words: one 32-byte store (rows <= sn are never read)
          if (r + 1 > sn) ll32_store2(reinterpret_cast<LL32*>(myrec + r), x[c].x, x[c].y, stamp);
          if (r == sn) al = x[c].x;
          if (r + 1 == sn) al = x[c].y;
          if (r > sn) ss = fma(x[c].x, x[c].x, ss);
          if (r + 1 > sn && r + 1 < l) ss = fma(x[c].y, x[c].y, ss);
        }
      }
      if (sn >= l) return;
#ifdef BRA_QRCP_TRACE
      if (p.ts && lane == 0 && sn == p.ts_step) p.ts[((size_t)cta * QR_WARPS + warp) * 16 + 14] = clock64();
#endif
      // row sn sits in exactly one lane
      const double alpha = __shfl_sync(0xffffffffu, al, ((sn - rb) & 63) >> 1);
      ss = warp_sum(ss);
#ifdef BRA_QRCP_TRACE
      if (p.ts && lane == 0 && sn == p.ts_step) p.ts[((size_t)cta * QR_WARPS + warp) * 16 + 15] = clock64();
#endif
      double beta, tau, scale;
      if (sn >= l - 1 || ss == 0.0) {
        beta = alpha;
        tau = 0.0;
        scale = 0.0;        // v = e_1
      } else {
        beta = -copysign(sqrt(fma(alpha, alpha, ss)), alpha);
        tau = (beta - alpha) / beta;
        scale = 1.0 / (alpha - beta);
      }
      if (lane < 4) {
        const double hv = lane == 0 ? tau : lane == 1 ? beta : lane == 2 ? __longlong_as_double((long long)(uint32_t)(col0 + cand_lc)) : scale;
        ll_store_d(myrec - RECH + lane, hv, stamp);
      }
    };

    // ---- pass 2 of step sp: a_j -= f_j v on every live column that still carries a non-zero f ----
    auto pass2 = [&](auto tag, const int sp, const unsigned live) {
      constexpr int NCH = decltype(tag)::value;
      const int i0 = sp >> 6;
      const int rbase = (i0 << 6) + 2 * lane;
      double2 vr[NCH];
      {
        const double2* v2 = reinterpret_cast<const double2*>(vbuf + (sp & 1) * LV + rbase);
#pragma unroll
        for (int c = 0; c < NCH; ++c) vr[c] = v2[32 * c];
      }
      double fj[JW];
#pragma unroll
      for (int j = 0; j < JW; ++j) fj[j] = ((live >> j) & 1) ? fbuf[warp + QF_CW * j] : 0.0;
#pragma unroll
      for (int j = JW - 1; j >= 0; --j) {
        const double f = fj[j];
        if (f == 0.0) continue;               // warp-uniform: dead, flagged (done in pass 1) or the candidate (done)
        const int cls = cls_of(j);
        double2 x[NCH];
        load_col(tag, j, cls, i0, x);
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
          x[c].x = fma(-f, vr[c].x, x[c].x);
          x[c].y = fma(-f, vr[c].y, x[c].y);
        }
        store_col(tag, j, cls, i0, x);
      }
      tm_wait_st();
    };

    // candidates for step 0: the initial norms themselves (step-0 keys are norms, later keys squared norms: keys are
    // only ever compared within one step)
    {
      long long key = QF_NONE;
      int lp = 0x7fffffff, ps = -1;
      if (mycol) {
        lp = lpos[mylc];
        key = p.nopivot ? (lp == 0 ? __double_as_longlong(1.0) : QF_NONE) : __double_as_longlong(sv1[mylc]);
        if (lp == 0) ps = (int)(col0 + mylc);
      }
      publish(key, lp, ps, 0);
    }
    qf_bar(2);

    bool pend2 = false;
    while (true) {
      QF_TS(4)
      // ---- am I the owner of this CTA's candidate for step s?  (my candidate beats the 14 others) ----
      {
        const QfSlotA mine = slotA[warp];
        const int cand_lc = slotB[warp].lc;
        bool better = false;
        if (lane < QF_CW && lane != warp) {
          const QfSlotA o = slotA[lane];
          better = (o.key > mine.key) || (o.key == mine.key && o.lp < mine.lp);
        }
        const bool owner = cand_lc >= 0 && !__any_sync(0xffffffffu, better);
        const int sp = s - 1;
        const int nchv = nchtot - ((sp >= 0 ? sp : 0) >> 6);
        if (owner) {
          QF_DISPATCH(nchv, cand_dlarfg, sp, cand_lc)
          tm_wait_st();
        }
        QF_TS(5)
        if (pend2) {
          const double temp = mycol ? stmp[mylc] : -1.0;
          if (temp >= 0.0) refresh(temp);
          const unsigned live = slive[warp];
          QF_DISPATCH(nchv, pass2, sp, live)
        }
        pend2 = false;
      }
      QF_TS(6)
#ifdef BRA_QRCP_TRACE
      if (warp == 0 && lane == 0) QF_TICK(1)
#endif
      qf_bar(3);                  // the comm warp knows the winner of step s
      if (ctrl->pre) break;
      if (warp < nchtot - (s >> 6)) fetch_chunk(s, ctrl->wcta, warp);
      qf_bar(1);                  // the record of step s is in vbuf; tau, lpos are set
      if (ctrl->fail) break;
      const int stop = ctrl->stop;
      const double tau = ctrl->tau;
      QF_TS(7)
      const bool downdate = (s < lastrk - 1) && !p.nopivot;      // no pivoting: the norms are never looked at
      {
        const int nchv = nchtot - (s >> 6);
        QF_DISPATCH(nchv, pass1, s, tau, downdate, stop == 0)
      }
      pend2 = true;
      QF_TS(8)
#ifdef BRA_QRCP_TRACE
      if (warp == 0 && lane == 0) QF_TICK(4)
#endif
      if (stop == 1) {
        // the reflector of the last executed step is applied in full (B keeps every update to the block end)
        const int nchv = nchtot - (s >> 6);
        const unsigned live = slive[warp];
        QF_DISPATCH(nchv, pass2, s, live)
        break;
      }
      ++s;
      qf_bar(2);                  // candidates for step s are published
    }
  }

  // ---- epilogue: the slab goes back to global memory (every warp its own columns), jpvt is completed ----
  __syncthreads();
  const int nsteps = ctrl->nsteps;     // pivoted columns
  if (warp < QF_CW) {
    for (int j = 0; warp + QF_CW * j < ncols; ++j) {
      const int lc = warp + QF_CW * j;
      double* g = p.B + (col0 + lc) * p.ldb;
      int slot;
      const int cls = class_of(warp, j, slot);
      if (cls == 0) {
        // a pivoted column was stored from its pivot row down when it was chosen (fetch_chunk): only R above it is here
        const int lp = lpos[lc];
        const int rlim = lp < nsteps ? lp : l;
        for (int c = 0; c < nchtot && 64 * c < rlim; ++c) {
          double2 x;
          tm_ld2(tm_col(warp, j) + 4 * c, x);
          tm_wait_ld();
          const int r = 64 * c + 2 * lane;
          if (r < rlim) g[r] = x.x;
          if (r + 1 < rlim) g[r + 1] = x.y;
        }
      } else if (cls == 1) {
        const double* d = cache + (size_t)slot * lds;
        for (int r = lane; r < l; r += 32) g[r] = d[r];
      }
    }
  }
  for (int lc = tid; lc < ncols; lc += QF_THREADS) {
    int lp = lpos[lc];
    if (lp >= nsteps) p.jpvt[lp] = col0 + lc + 1;
  }
  if (cta == 0 && tid == 0) {
    p.info[1] = nsteps;
#ifdef BRA_QRCP_TRACE
    for (int i = 0; i < 5; ++i) p.info[4 + i] = (int)(s_tph[i] >> 10);   // kilo-cycles per phase
#else
    for (int i = 0; i < 5; ++i) p.info[4 + i] = 0;
#endif
  }
#ifdef BRA_QRCP_TRACE
  if (tid == 0 && p.dbg)
    for (int i = 0; i < 5; ++i) p.dbg[cta * 8 + i] = (int)(s_tph[i] >> 10);
#endif
  __syncthreads();
  if (warp == QF_CW) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tbase) : "memory");
#undef vbuf
#undef ctrl
#undef rdblk
#undef slotA
#undef slotB
#undef fbuf
#undef st2
#undef ssq
#undef sv1
#undef srv2
#undef stmp
#undef slive
#undef lpos
#undef cache
}

size_t fast_fixed_bytes(int l, int cpc, int nbe) {
  const int LV = ((l + 63) >> 6) << 6;
  const int cpe = (cpc + 1) & ~1;
  return (size_t)2 * LV * 8 + 32 + (size_t)((nbe + 1) & ~1) * 8 + 16 * 16 + 16 * 8 + 16 * 4 + (size_t)cpe * 8 * 7 +
         (size_t)((cpe + 3) & ~3) * 4 + 64;
}

}  // namespace

// Can the fast kernel take this shape?  Fills the column-per-warp variant, the slab stride, the number of columns that
// fit in shared memory and the dynamic shared-memory size.
bool bra_qrcp_fast_plan(int l, int cpc, int nbe, size_t budget, bool aligned, int* jw, int* lds, int* csm, size_t* smem) {
  if (l > 64 * QF_MAXCH || cpc > QF_CW * 8) return false;
  const int ld = (l + 1) & ~1;
  const int nch = (l + 63) >> 6;
  const size_t fixed = fast_fixed_bytes(l, cpc, nbe);
  if (fixed > budget) return false;
  // columns per warp beyond the tensor-memory capacity of the tightest quadrant go to shared-memory slots
  // ((j - tcap) * 15 + warp), what does not fit there stays in global memory (L2)
  static const bool no_tmem = getenv("BRA_QRCP_NOTMEM") != nullptr;
  const int cpw = (cpc + QF_CW - 1) / QF_CW;
  int tcap_min = 1 << 30;
  for (int w = 0; w < QF_CW; ++w) tcap_min = std::min(tcap_min, qf_tcap(w, 4 * nch));
  if (no_tmem) tcap_min = 0;
  int want = cpw > tcap_min ? (cpw - tcap_min) * QF_CW : 0;
  int c = (int)((budget - fixed) / ((size_t)ld * 8));
  if (c > want) c = want;
  if (c < want && !aligned) return false;      // L2-resident columns need 16-byte aligned 128-bit accesses
  *jw = cpc <= QF_CW * 4 ? 4 : 8;
  *lds = ld;
  *csm = c;
  *smem = fixed + (size_t)c * ld * 8;
  return true;
}

template <int JW, int NMAX>
static cudaError_t fast_launch(const QrcpParams& p, int G, size_t smem, cudaStream_t st) {
  void* args[] = {(void*)&p};
  cudaError_t e = cudaFuncSetAttribute(qrcp_fast_kernel<JW, NMAX>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  return cudaLaunchCooperativeKernel((void*)qrcp_fast_kernel<JW, NMAX>, dim3(G), dim3(QF_THREADS), args, smem, st);
}

cudaError_t bra_qrcp_fast_launch(const QrcpParams& p, int G, int jw, size_t smem, cudaStream_t st) {
  const int nch = (p.l + 63) >> 6;
  if (jw == 4) {
    if (nch <= 3) return fast_launch<4, 3>(p, G, smem, st);
    if (nch <= 5) return fast_launch<4, 5>(p, G, smem, st);
    return fast_launch<4, 9>(p, G, smem, st);
  }
  if (nch <= 3) return fast_launch<8, 3>(p, G, smem, st);
  if (nch <= 5) return fast_launch<8, 5>(p, G, smem, st);
  return fast_launch<8, 9>(p, G, smem, st);
}

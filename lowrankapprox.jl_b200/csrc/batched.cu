// Batched idfact of many independent small blocks (BASELINE config 5: 16384 HODLR-style 512 x 512 off-diagonal
// blocks, sketch = :sprn).  The reference has no batched API: this is the loop
//     for b in blocks: idfact(A_b; sketch=:sprn, ...)          src/id.jl:434-447 -> src/sketch.jl:674-690
// fused into ONE kernel, one CTA per block at a time:
//   sparse-Gaussian sketch (src/sketch.jl:571-589)  ->  early-terminating QRCP with dlaqps semantics
//   (src/pqr.jl:361-418)  ->  T = R11^{-1} R12 (src/pqr.jl:438-442),
// with the l x n sketch living entirely on chip: A_b streams through shared memory in 16-column tiles (cp.async,
// double buffered, each entry read from HBM exactly once, coalesced; stored row-major with an odd stride so that the
// permuted row gather of the sparse sketch is bank-conflict free), then one COLUMN PER THREAD in
// registers for the factorization, so that the Householder dot products and rank-1 updates need no cross-lane
// traffic at all; only the pivot search (redux.sync argmax + 16-entry merge) and the pivot column broadcast go
// through shared memory.  HBM-bound by design: 8 m n bytes per block in, p, k and the k x (n-k) T out.
//
// This kernel covers the first adaptive round (order = nb <= 32, n <= 512, m*8*16 <= 64 KB of column stages);
// blocks that do not terminate in it (k >= nb) and other shapes go through the general single-matrix path.
#include "common.cuh"
#include "qrcp_common.cuh"
#include <algorithm>
#include <cstdlib>
#include <vector>

namespace {

constexpr int BT = 512;            // threads = columns per block
constexpr int BW = BT / 32;        // warps
constexpr int BL = 32;             // sketch rows held in registers per thread

struct BatchParams {
  const double* A;
  int64_t lda, strideA;
  int m, n, l, kcap, nb;
  int nstages;             // tile stages in shared memory (3, fewer when the block is too tall)
  double atol, rtol;
  const int64_t* perm;     // 1-based randperm(m); block b at perm + b*perm_stride (0 = shared by all blocks);
                           // nullptr: fast mode -- every block draws its OWN permutation and weights in the kernel
                           // (the reference draws per call of sketch_sprn, i.e. per block: src/sketch.jl:575-579)
  uint64_t seed;           // fast mode: key of the per-block streams
  int64_t perm_stride;
  const double* s;         // weights, same layout
  int64_t s_stride;
  const int32_t* blocks;   // optional list of block ids to process (nullptr: 0..nblocks-1)
  int nblocks;
  int64_t* kout;           // [block id]
  int64_t* pout;           // [block id][n], 1-based
  double* Tout;            // block b: k_b x (n - k_b) at Tout + b*strideT, leading dimension ldT
  int64_t ldT, strideT;
  int32_t* status;         // [block id]: 0 ok, 1 T slot too small (k_b > ldT)
  long long* dbg;          // [CTA][8] cycles: tables, waiting for tiles, QRCP, outputs+T, tile issue, gather, hand-over to registers
};

__device__ __forceinline__ uint64_t bsplitmix(uint64_t s0, uint64_t i) {      // (i+1)-th output of SplitMix64 seeded s0
  uint64_t z = s0 + (i + 1) * 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

__device__ __forceinline__ uint32_t bsmem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void bmbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  const long long t0 = clock64();
  do {
    if (clock64() - t0 > (1ll << 31)) break;        // ~1 s: never hang the device (the results then fail their checks)
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(bsmem_u32(bar)), "r"(parity)
        : "memory");
  } while (!ok);
}

// BULK: the tiles arrive as one bulk copy per column (the copy engine, 4 KB each, completion on an mbarrier) into a
// column-major stage [16][CS]; otherwise as 8-byte cp.async into a row-major stage [m][17] (any alignment).  Measured
// on B200: a warp-wide 8-byte cp.async with 32 scattered shared-memory targets costs ~16 LSU cycles, i.e. 16 B/clk per
// SM -- the whole sketch phase ran at that rate.
template <bool BULK>
__global__ void __launch_bounds__(BT, 1) idfact_batched_kernel(BatchParams P) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ __align__(8) uint64_t mbar[3];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int m = P.m, n = P.n, l = P.l;
  constexpr int TC = 16;                                         // columns of A_b per tile
  constexpr int TS = TC + 1;                                     // row stride of a tile (odd: conflict-free both ways)
  // BULK: column stride = 2 (mod 4) doubles: 16-byte aligned columns, and the 16 columns of one row fall on 8 distinct
  // 8-byte banks (a 2-way conflict in the gather instead of 16-way)
  const int CS = (m % 4 == 2) ? m : m + 2;
  const int rstr = BULK ? 1 : TS, cstr = BULK ? CS : 1;           // element (r, c) of a stage at r * rstr + c * cstr
  uint32_t mphase = 0;                                           // BULK: phase bit of each stage's barrier
  if (BULK) {
    if (tid == 0) {
      for (int i = 0; i < 3; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bsmem_u32(&mbar[i])));
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
  }
  const int nst = P.nstages;                                     // 3 / 2: tiles in flight behind the current one; 1: tall blocks
  const size_t tile_elems = (size_t)max(m, 2 * BL) * TS;
  double* tiles = reinterpret_cast<double*>(smem_raw);          // [nst][m][TS]   (later: R11, [BL][BL+1])
  double* Xb = tiles + (size_t)nst * tile_elems;                 // [BL][TS] sketch of the current tile
  const int tpad = (((m / l) + 1) * l + 1) & ~1;                 // table entries, [t][i] layout
  double* tabw = Xb + (size_t)BL * TS;                           // [tpad] weight of term t of sketch row i at t*l + i
  int* tabo = reinterpret_cast<int*>(tabw + tpad);               // [tpad] offset of its (permuted) row in a stage
  double* vv = tabw + 2 * tpad;                                  // [BL] Householder vector
  double* rdblk = vv + BL;                                      // [BL] diagonal of R within the current block
  double* cv = rdblk + BL;                                      // [BW] warp candidates: norm
  double* hh = cv + BW;                                         // tau, beta
  int* clp = reinterpret_cast<int*>(hh + 2);                    // [BW] logical position
  int* ctd = clp + BW;                                          // [BW] owning thread
  double* colbuf = tiles;

  const int lmin = min(l, n);
  const int64_t q = m / l, rem = m % l;                          // p_i = q + (i < rem), off_i = i*q + min(i, rem)
  bool tables_loaded = false;
  long long tph[8] = {0, 0, 0, 0, 0, 0, 0, 0}, tlast = clock64();
#define BTICK(i) { long long _t = clock64(); tph[i] += _t - tlast; tlast = _t; }

  for (int it = blockIdx.x; it < P.nblocks; it += gridDim.x) {
    const int b = P.blocks ? P.blocks[it] : it;
    const double* Ab = P.A + (int64_t)b * P.strideA;
    if (BULK) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // R11 / tables were written by threads
    __syncthreads();
    // tile T = columns 16T..16T+15 of A_b, all m rows, staged as tile[r][c] (row stride 17) by 8-byte cp.async: warp w
    // copies column w (lanes = consecutive rows: coalesced global reads, conflict-free shared stores)
    auto issue = [&](int T) {
      if (BULK) {
        // one bulk copy per column, issued by lane 0 of warp c (a single thread issues them ~100 cycles apart); thread 0
        // posts the byte count -- a copy that completes first just drives the pending-byte count negative for a while
        const int nc = min(TC, n - T * TC);
        uint64_t* bar = &mbar[T % nst];
        if (tid == 0)
          asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bsmem_u32(bar)), "r"((uint32_t)(nc * m * 8))
                       : "memory");
        if (lane == 0 && warp < nc)
          asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                           bsmem_u32(tiles + (size_t)(T % nst) * tile_elems + (size_t)warp * CS)),
                       "l"(Ab + (int64_t)(T * TC + warp) * P.lda), "r"((uint32_t)(m * 8)), "r"(bsmem_u32(bar))
                       : "memory");
        return;
      }
      const int c = T * TC + warp;                 // BW == TC: one column per warp
      if (c < n) {
        const double* g = Ab + (int64_t)c * P.lda;
        double* d = tiles + (size_t)(T % nst) * tile_elems + warp;
        for (int r = lane; r < m; r += 32)
          asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((uint32_t)__cvta_generic_to_shared(d + (size_t)r * TS)),
                       "l"(g + r)
                       : "memory");
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    };
    // double buffered: the first tile is already on its way while the tables (and, in fast mode, this block's own
    // permutation) are built in the second stage's memory
    const bool early = nst > 1;
    if (early) issue(0);
    if (!tables_loaded || P.perm_stride != 0 || P.s_stride != 0 || P.perm == nullptr) {
      // tables in [t][i] order (term t of sketch row i at t*l + i): lanes = sketch rows read consecutive words
      const int64_t* pb = P.perm ? P.perm + (int64_t)b * P.perm_stride : nullptr;
      const double* sb = P.perm ? P.s + (int64_t)b * P.s_stride : nullptr;
      unsigned long long* skey = reinterpret_cast<unsigned long long*>(tiles + (early ? tile_elems : 0));   // a free stage
      const uint64_t bkey = P.seed * 0x9E3779B97F4A7C15ull + (uint64_t)(b + 1) * 0xD1B54A32D192ED03ull;
      if (!P.perm) {
        // randperm(m) for THIS block: m random 53-bit keys with the row index in the low 11 bits, bitonic sort in
        // shared memory (m <= ~1500 < 2048), the sorted indices are the permutation
        int n2 = 1;
        while (n2 < m) n2 <<= 1;
        if (n2 == BT) {
          // one key per thread (the C5 shape, 256 < m <= 512): bitonic network with the intra-warp stages on shuffles,
          // 32-bit keys (21 random bits | 11 index bits: ties are broken by the index, still a permutation)
          unsigned* sk32 = reinterpret_cast<unsigned*>(skey) + 2 * BT;        // exchange buffer behind the result
          unsigned v = tid < m ? (((unsigned)(bsplitmix(bkey, (uint64_t)tid) >> 32) & ~0x7FFu) | (unsigned)tid) : 0xFFFFFFFFu;
          for (int kk = 2; kk <= BT; kk <<= 1)
            for (int j = kk >> 1; j > 0; j >>= 1) {
              unsigned o;
              if (j >= 32) {
                sk32[tid] = v;
                __syncthreads();
                o = sk32[tid ^ j];
                __syncthreads();
              } else {
                o = __shfl_xor_sync(0xffffffffu, v, j);
              }
              const bool keep_min = ((tid & j) == 0) == ((tid & kk) == 0);
              v = keep_min ? min(v, o) : max(v, o);
            }
          skey[tid] = (unsigned long long)v;
          __syncthreads();
        } else {
        for (int r = tid; r < n2; r += BT)
          skey[r] = r < m ? (((bsplitmix(bkey, (uint64_t)r) >> 32) & ~0x7FFull) | (unsigned long long)r) : ~0ull;   // same keys as above
        __syncthreads();
        for (int kk = 2; kk <= n2; kk <<= 1)
          for (int j = kk >> 1; j > 0; j >>= 1) {
            for (int i = tid; i < n2; i += BT) {
              const int ixj = i ^ j;
              if (ixj > i) {
                const unsigned long long a0 = skey[i], a1 = skey[ixj];
                if (((i & kk) == 0) ? (a0 > a1) : (a0 < a1)) {
                  skey[i] = a1;
                  skey[ixj] = a0;
                }
              }
            }
            __syncthreads();
          }
        }
      }
      for (int r = tid; r < m; r += BT) {
        // r = off_i + t  ->  (i, t)
        int i, t;
        const int q32 = (int)q, rem32 = (int)rem;          // m < 2^31: 32-bit divisions
        if (r < rem32 * (q32 + 1)) {
          i = r / (q32 + 1);
          t = r - i * (q32 + 1);
        } else {
          const int r2 = r - rem32 * (q32 + 1);
          i = rem32 + r2 / q32;
          t = r2 % q32;
        }
        double wgt;
        long long prow;
        if (P.perm) {
          wgt = sb[r];
          prow = pb[r] - 1;
        } else {
          // N(0,1) weight of term r of this block: Box-Muller on two counter-based uniforms
          const uint64_t z1 = bsplitmix(bkey ^ 0xA5A5A5A5A5A5A5A5ull, (uint64_t)(2 * r));
          const uint64_t z2 = bsplitmix(bkey ^ 0xA5A5A5A5A5A5A5A5ull, (uint64_t)(2 * r + 1));
          const double u1 = ((double)(z1 >> 11) + 1.0) * (1.0 / 9007199254740992.0);
          const double u2 = ((double)(z2 >> 11) + 0.5) * (1.0 / 9007199254740992.0);
          double sn, cs;
          sincospi(2.0 * u2, &sn, &cs);
          wgt = sqrt(-2.0 * log(u1)) * cs;
          prow = (long long)(skey[r] & 0x7FFull);
        }
        tabw[t * l + i] = wgt;
        tabo[t * l + i] = (int)prow * rstr;
      }
      tables_loaded = true;
      if (BULK) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // the key buffers become a tile stage
      __syncthreads();
    }

    BTICK(0)
    // ---- sparse-Gaussian sketch, tile by tile.  Tile T = columns 16T..16T+15 of A_b, all m rows, staged as
    // tile[r][c] (row stride 17) by 8-byte cp.async: warp w copies column w (lanes = consecutive rows: coalesced
    // global reads, conflict-free shared stores).  Thread (i = tid / 16, c = tid % 16) then forms B[i, 16T + c] in the
    // reference's summation order; its 16 terms read tile[perm[t, i]][c] -- one permuted ROW, contiguous across the
    // half-warp -- so the gather costs no bank conflicts.  The finished 32 x 16 piece goes through a small exchange
    // buffer straight into the registers of the 16 threads that own those columns for the factorization.
    const bool live = tid < n;
    double a[BL];
#pragma unroll
    for (int i = 0; i < BL; ++i) a[i] = 0.0;
    {
      const int ntiles = (n + TC - 1) / TC;
      if (!early) issue(0);
      // sketch row / column within the tile.  A half-warp (one wavefront of an 8-byte load) takes 8 columns of TWO
      // sketch rows: in the column-major BULK stage the 16 columns of one row sit on only 8 distinct banks, two rows of
      // different parity fill all 16; the weight read (8 bytes, two addresses per half-warp) is one wavefront per half,
      // the offset read (4 bytes) one per warp -- 3 + 3 wavefronts per term instead of 4 + 4
      const int si = 2 * (tid >> 5) + ((tid >> 3) & 1), sc = (tid & 7) | ((tid >> 1) & 8);
      const int pi = (si < l) ? (int)q + (si < rem ? 1 : 0) : 0;
      if (nst > 2 && ntiles > 1) issue(1);           // three stages: two tiles in flight behind the one being read
      for (int T = 0; T < ntiles; ++T) {
        // tile T + nst - 1 goes into the stage that was last read in iteration T-1 (barrier below); then wait until at
        // most the tiles after T are still pending
        if (nst > 1 && T + nst - 1 < ntiles) issue(T + nst - 1);
        const int later = min(ntiles - 1 - T, nst - 1);
        BTICK(4)
        if (BULK) {
          bmbar_wait(&mbar[T % nst], (mphase >> (T % nst)) & 1u);
          mphase ^= 1u << (T % nst);
        } else if (later >= 2) asm volatile("cp.async.wait_group 2;" ::: "memory");
        else if (later == 1) asm volatile("cp.async.wait_group 1;" ::: "memory");
        else asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();
        BTICK(1)
        const double* tl = tiles + (size_t)(T % nst) * tile_elems;
        if (si < l) {
          double acc = 0.0;
#pragma unroll 4
          for (int t = 0; t < pi; ++t) {
            acc += tabw[t * l + si] * tl[tabo[t * l + si] + sc * cstr];
          }
          Xb[si * TS + sc] = acc;
        }
        __syncthreads();
        BTICK(5)
        if (nst == 1 && T + 1 < ntiles) issue(T + 1);
        if ((tid >> 4) == T && live) {
#pragma unroll
          for (int i = 0; i < BL; ++i)
            if (i < l) a[i] = Xb[i * TS + (tid & 15)];
        }
        BTICK(6)
      }
    }
    __syncthreads();
    BTICK(6)

    // ---- one column per thread, in registers ----
    double vn1, vn2;
    {
      // initial norm (src/pqr.jl:376-385), power-of-two scaled like the grid-wide kernel
      double amax = 0.0;
#pragma unroll
      for (int i = 0; i < BL; ++i) amax = fmax(amax, fabs(a[i]));
      double nrm = 0.0;
      if (amax > 0.0) {
        const int e = ilogb(amax);
        const double sc = scalbn(1.0, -e);
        double ss = 0.0;
#pragma unroll
        for (int i = 0; i < BL; ++i) {
          const double x = a[i] * sc;
          ss = fma(x, x, ss);
        }
        nrm = scalbn(sqrt(ss), e);
      }
      vn1 = vn2 = nrm;
    }
    int lpos = live ? tid : 0x7fffffff;

    int s = 0, jblk = 0, cnt = 0, jb = min(P.nb, P.kcap), kres = (P.kcap == 0) ? 0 : -1;
    double ptol = 0.0;
    int pend_flag = 0;          // a column was flagged in the previous step
    while (kres < 0) {
      // ---- pivot search: first maximum of the downdated norms (idamax) ----
      {
        const double v = (live && lpos >= s) ? vn1 : -1.0;
        const int wl = warp_argmax(v, lpos);
        if (lane == wl) {
          cv[warp] = v;
          clp[warp] = lpos;
          ctd[warp] = tid;
        }
      }
      __syncthreads();
      double wv;
      int wlp, wt;
      {
        const double v = (lane < BW) ? cv[lane] : -1.0;
        const int lp = (lane < BW) ? clp[lane] : 0x7fffffff;
        const int wl = warp_argmax(v, lp);
        wv = __shfl_sync(0xffffffffu, v, wl);
        wlp = __shfl_sync(0xffffffffu, lp, wl);
        wt = ctd[wl];
      }
      // ---- block bookkeeping for the previous step (a flagged column ends dlaqps' block) ----
      if (cnt > 0 && pend_flag) {
        if (fabs(rdblk[cnt - 1]) <= ptol) {
          for (int i = 0; i < cnt; ++i)
            if (fabs(rdblk[i]) <= ptol) {
              kres = jblk + i;
              break;
            }
        }
        if (kres >= 0) break;
        jblk += cnt;
        cnt = 0;
        jb = min(P.nb, P.kcap - jblk);
      }
      if (s == 0) ptol = fmax(P.atol, P.rtol * wv);                  // src/pqr.jl:386-389

      // ---- pivot column -> shared memory; warp 0 runs dlarfg on it ----
      if (tid == wt) {
#pragma unroll
        for (int i = 0; i < BL; ++i) vv[i] = a[i];
      }
      __syncthreads();
      if (warp == 0) {
        const double x = (lane > s && lane < l) ? vv[lane] : 0.0;
        const double ssq = warp_sum(x * x);
        const double alpha = vv[s];
        double beta, tau, scale;
        if (s >= l - 1 || ssq == 0.0) {
          beta = alpha;
          tau = 0.0;
          scale = 0.0;
        } else {
          beta = -copysign(sqrt(fma(alpha, alpha, ssq)), alpha);
          tau = (beta - alpha) / beta;
          scale = 1.0 / (alpha - beta);
        }
        __syncwarp();
        vv[lane] = (lane == s) ? 1.0 : x * scale;                   // v: 1 at row s, zeros above, scaled tail below
        if (lane == 0) {
          hh[0] = tau;
          hh[1] = beta;
          rdblk[cnt] = beta;
        }
      }
      __syncthreads();
      const double tau = hh[0], beta = hh[1];
      // ---- ownership (no physical swaps: logical positions move instead) ----
      if (live && lpos == s && tid != wt) lpos = wlp;                // column K takes the pivot's old position
      if (tid == wt) {
        lpos = s;
#pragma unroll
        for (int i = 0; i < BL; ++i)
          if (i == s) a[i] = beta;                                   // R[s, s]
      }
      // ---- apply H to my column, downdate its norm (dlaqps steps 5-8 fused) ----
      int flagged = 0;
      if (live && lpos > s) {
        double dot = 0.0;
#pragma unroll
        for (int i = 0; i < BL; ++i)
          if (i >= s) dot = fma(a[i], vv[i], dot);
        const double f = tau * dot;
        double rs = 0.0;
#pragma unroll
        for (int i = 0; i < BL; ++i)
          if (i >= s) {
            a[i] = fma(-f, vv[i], a[i]);
            if (i == s) rs = a[i];
          }
        if (s < lmin - 1 && vn1 != 0.0) {
          double t = fabs(rs) / vn1;
          t = fmax(0.0, (1.0 + t) * (1.0 - t));
          const double r2 = vn1 / vn2;
          const double t2 = t * (r2 * r2);
          if (t2 <= TOL3Z) {
            // flagged: norm recomputed from rows s+1.. of the updated column
            double ss = 0.0;
#pragma unroll
            for (int i = 0; i < BL; ++i)
              if (i > s) ss = fma(a[i], a[i], ss);
            vn1 = vn2 = sqrt(ss);
            flagged = 1;
          } else {
            vn1 *= sqrt(t);
          }
        }
      }
      pend_flag = __syncthreads_or(flagged);
      // ---- end of step ----
      ++cnt;
      ++s;
      if (cnt == jb) {
        // block ends by count; flags raised in this step are irrelevant
        pend_flag = 0;
        if (fabs(rdblk[cnt - 1]) <= ptol) {
          for (int i = 0; i < cnt; ++i)
            if (fabs(rdblk[i]) <= ptol) {
              kres = jblk + i;
              break;
            }
        }
        if (kres >= 0) break;
        jblk += cnt;
        cnt = 0;
        if (jblk >= P.kcap) {
          kres = P.kcap;
          break;
        }
        jb = min(P.nb, P.kcap - jblk);
      }
    }
    const int k = kres;
    BTICK(2)

    // ---- outputs: p, k, T = R11^{-1} R12 ----
    if (live) P.pout[(int64_t)b * n + lpos] = (int64_t)tid + 1;
    if (tid == 0) {
      P.kout[b] = k;
      P.status[b] = (k > P.ldT) ? 1 : 0;
    }
    __syncthreads();
    double* R11 = colbuf;                                            // [BL][BL+1], row-major with padding
    if (live && lpos < k) {
#pragma unroll
      for (int i = 0; i < BL; ++i)
        if (i < k) R11[i * (BL + 1) + lpos] = (i <= lpos) ? a[i] : 0.0;
    }
    __syncthreads();
    if (live && lpos >= k && k <= P.ldT) {
      // back substitution on my column of R12 (dtrsm L,U,N,N), entirely in registers, column-oriented: once x_i is
      // known it is eliminated from all rows above (independent FMAs instead of one dependent dot product per row)
#pragma unroll
      for (int i = BL - 1; i >= 0; --i) {
        if (i < k) {
          const double x = a[i] / R11[i * (BL + 1) + i];
          a[i] = x;
#pragma unroll
          for (int j = 0; j < i; ++j) a[j] = fma(-R11[j * (BL + 1) + i], x, a[j]);
        }
      }
      double* t = P.Tout + (int64_t)b * P.strideT + (int64_t)(lpos - k) * P.ldT;
#pragma unroll
      for (int i = 0; i < BL; ++i)
        if (i < k) t[i] = a[i];
    }
    BTICK(3)
  }
  if (tid == 0 && P.dbg)
    for (int i = 0; i < 8; ++i) P.dbg[blockIdx.x * 8 + i] = tph[i];
#undef BTICK
}

}  // namespace

int bra_fill_meta(bra_ctx* ctx, int kind, void* dst_dev, int64_t count, int64_t range, uint64_t seed, uint64_t stream_id);

extern "C" {

int bra_idfact_batched_f64(bra_ctx* ctx, int64_t nblocks, int64_t m, int64_t n, const double* A, int64_t lda,
                           int64_t strideA, const bra_opts* opts, const int64_t* perm, int64_t perm_stride,
                           const double* s, int64_t s_stride, int64_t* k_out, int64_t* p_out, double* T_out,
                           int64_t ldT, int64_t strideT) {
  if (!ctx) return -1;
  BRA_CHECK_ARG(nblocks >= 0, 2, "nblocks");
  BRA_CHECK_ARG(m >= 1, 3, "m");
  BRA_CHECK_ARG(n >= 1, 4, "n");
  BRA_CHECK_ARG(A != nullptr || nblocks == 0, 5, "A");
  BRA_CHECK_ARG(lda >= m, 6, "lda");
  BRA_CHECK_ARG(strideA >= lda * (n - 1) + m || nblocks <= 1, 7, "strideA");
  if (bra_chkopts(ctx, opts)) return -8;
  BRA_CHECK_ARG((perm == nullptr) == (s == nullptr), 9, "perm and s must be given together");
  BRA_CHECK_ARG(k_out != nullptr && p_out != nullptr && T_out != nullptr, 13, "outputs");
  BRA_CHECK_ARG(ldT >= 1 && strideT >= ldT * n, 16, "ldT/strideT");
  if (nblocks == 0) return BRA_OK;
  BRA_CUDA(cudaSetDevice(ctx->device));
  if (!is_device_ptr(A) || !is_device_ptr(k_out) || !is_device_ptr(p_out) || !is_device_ptr(T_out) ||
      (perm && (!is_device_ptr(perm) || !is_device_ptr(s)))) {
    ctx->set_error("bra_idfact_batched_f64 takes device-resident blocks, random inputs and outputs");
    return BRA_ERR_UNSUPPORTED;
  }
  if (opts->sketch != BRA_SKETCH_SPRN || opts->maxdet_tol >= 0 || opts->sketch_randn_niter > 0) {
    ctx->set_error("the batched kernel is built for sketch = :sprn (BASELINE config 5)");
    return BRA_ERR_UNSUPPORTED;
  }
  // first adaptive round: order = nb (src/sketch.jl:677-680); non-adaptive: order = rank (:686)
  const bool adaptive = opts->sketchfact_adap || opts->rank < 0;
  const int64_t order = adaptive ? opts->nb : opts->rank;
  const int64_t ordc = order > 0 ? order : 1;
  const int64_t tpad = (((m / ordc) + 1) * ordc + 1) & ~int64_t(1);
  const int64_t tile_elems = (m > 2 * BL ? m : 2 * BL) * 17;
  auto smem_for = [&](int nst) {
    return ((size_t)nst * tile_elems + (size_t)BL * 17 + 2 * tpad + 2 * BL + BW + 2) * 8 + (2 * BW) * 4 + 64;
  };
  int nstages = 2;                  // measured on C5: a third stage buys nothing (the tile wait is 4 % of a block); tall blocks: one
  if (const char* ev = getenv("BRA_BATCHED_STAGES")) nstages = std::max(1, std::min(3, atoi(ev)));
  while (nstages > 1 && smem_for(nstages) > (size_t)ctx->smem_optin) --nstages;
  const size_t smem = smem_for(nstages);
  if (order < 1 || order > BL || n > BT || order > m || smem > (size_t)ctx->smem_optin) {
    ctx->set_error("batched idfact: shape outside the fused kernel (needs order = nb <= 32, n <= 512, "
                   "32 <= m <= ~1500); factor such blocks one by one with bra_idfact_f64");
    return BRA_ERR_UNSUPPORTED;
  }
  const int64_t lmin = order < n ? order : n;
  const int64_t kcap = (opts->rank < 0 || opts->rank > lmin) ? lmin : opts->rank;

  BatchParams P;
  P.A = A;
  P.lda = lda;
  P.strideA = strideA;
  P.m = (int)m;
  P.n = (int)n;
  P.l = (int)order;
  P.kcap = (int)kcap;
  P.nstages = nstages;
  P.nb = (int)(opts->nb < kcap ? opts->nb : (kcap > 0 ? kcap : 1));
  P.atol = opts->atol;
  P.rtol = opts->rtol;
  if (perm) {
    P.perm = perm;
    P.perm_stride = perm_stride;
    P.s = s;
    P.s_stride = s_stride;
  } else {
    // fast mode: every block draws its own permutation and weights inside the kernel (keyed by opts.seed and the block id)
    P.perm = nullptr;
    P.perm_stride = 0;
    P.s = nullptr;
    P.s_stride = 0;
  }
  P.seed = opts->seed;
  P.blocks = nullptr;
  P.nblocks = (int)nblocks;
  P.kout = k_out;
  P.pout = p_out;
  P.Tout = T_out;
  P.ldT = ldT;
  P.strideT = strideT;
  BRA_CUDA(ctx->scratch.reserve((size_t)nblocks * 4));
  P.status = ctx->scratch.as<int32_t>();
  BRA_CUDA(ctx->scratch3.reserve((size_t)ctx->num_sms * 8 * 8));
  P.dbg = ctx->scratch3.as<long long>();
  // bulk copies need 16-byte aligned columns of 16-byte multiples
  const bool bulk = (m % 2 == 0) && (lda % 2 == 0) && (strideA % 2 == 0 || nblocks <= 1) &&
                    (reinterpret_cast<uintptr_t>(A) % 16 == 0) && getenv("BRA_BATCHED_NOBULK") == nullptr;
  const int grid = (int)std::min<int64_t>(nblocks, ctx->num_sms);
  if (bulk) {
    BRA_CUDA(cudaFuncSetAttribute(idfact_batched_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ProfScope ps(ctx, BRA_PROF_BATCHED);
    idfact_batched_kernel<true><<<grid, BT, smem, ctx->stream>>>(P);
  } else {
    BRA_CUDA(cudaFuncSetAttribute(idfact_batched_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ProfScope ps(ctx, BRA_PROF_BATCHED);
    idfact_batched_kernel<false><<<grid, BT, smem, ctx->stream>>>(P);
  }
  ctx->launches++;
  BRA_CUDA(cudaGetLastError());
  // k is data dependent: read it back (the only host sync), report blocks the fused round could not finish
  std::vector<int64_t> hk((size_t)nblocks);
  std::vector<int32_t> hs((size_t)nblocks);
  BRA_CUDA(cudaMemcpyAsync(hk.data(), k_out, (size_t)nblocks * 8, cudaMemcpyDeviceToHost, ctx->stream));
  BRA_CUDA(cudaMemcpyAsync(hs.data(), P.status, (size_t)nblocks * 4, cudaMemcpyDeviceToHost, ctx->stream));
  BRA_CUDA(cudaStreamSynchronize(ctx->stream));
  int64_t unfinished = 0, tslot = 0;
  for (int64_t b = 0; b < nblocks; ++b) {
    if (hs[(size_t)b]) ++tslot;
    // the adaptive loop continues while k >= n_t (src/sketch.jl:681-682): such blocks go through the general
    // single-matrix path from the next round on (fresh library-drawn random inputs keyed by seed, round, block)
    if (adaptive && hk[(size_t)b] >= order) {
      ++unfinished;
      bra_opts ob = *opts;
      ob.seed = opts->seed + 0x9E3779B97F4A7C15ull * (uint64_t)(b + 1);
      ctx->start_round = 1;
      int rc = bra_sketchfact_core(ctx, 'n', m, n, A + b * strideA, lda, &ob, nullptr);
      ctx->start_round = 0;
      if (rc) return rc;
      const int64_t kb = ctx->res.k;
      BRA_CUDA(cudaMemcpyAsync(k_out + b, &kb, 8, cudaMemcpyHostToDevice, ctx->stream));
      BRA_CUDA(cudaMemcpyAsync(p_out + b * n, ctx->jpvt.p, (size_t)n * 8, cudaMemcpyDeviceToDevice, ctx->stream));
      if (kb > ldT) {
        ++tslot;
      } else if (kb > 0 && n > kb) {
        BRA_CUDA(cudaMemcpy2DAsync(T_out + b * strideT, (size_t)ldT * 8, ctx->T.p, (size_t)ctx->res.ldT * 8, (size_t)kb * 8,
                                   (size_t)(n - kb), cudaMemcpyDeviceToDevice, ctx->stream));
      }
      BRA_CUDA(cudaStreamSynchronize(ctx->stream));
    }
  }
  ctx->batched_unfinished = unfinished;
  if (tslot) {
    ctx->set_error("batched idfact: " + std::to_string(tslot) + " blocks have k > ldT; enlarge the T slots");
    return BRA_ERR_TSLOT;
  }
  return BRA_OK;
}

int64_t bra_batched_unfinished(bra_ctx* ctx) { return ctx ? ctx->batched_unfinished : -1; }

int bra_debug_batched_phases(bra_ctx* ctx, int64_t* out8) {
  if (!ctx || !out8) return -1;
  long long h[8];
  BRA_CUDA(cudaMemcpy(h, ctx->scratch3.p, 64, cudaMemcpyDeviceToHost));
  for (int i = 0; i < 8; ++i) out8[i] = h[i];
  return BRA_OK;
}

}  // extern "C"

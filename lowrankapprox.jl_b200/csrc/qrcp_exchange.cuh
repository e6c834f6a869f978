// Exchange primitives and launch parameters shared by the persistent QRCP kernels (qrcp.cu: the general kernel,
// qrcp_fast.cu: the warp-specialised kernel for short sketches).  "LL" words are self-validating: every 4-byte
// payload is followed by the 4-byte step stamp (the NCCL low-latency idea), so a reader needs no fence or flag.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace braq {

constexpr int QR_THREADS = 512;
constexpr int QR_WARPS = QR_THREADS / 32;
constexpr int RECH = 4;                        // record header words: tau, beta, physical column, (pad)
constexpr int MAXG = 160;                      // >= number of SMs
constexpr uint32_t SPIN_LIMIT = 1u << 22;      // exchange timeout (never hang the box)

struct __align__(16) LL16 {
  uint32_t lo, s0, hi, s1;
};

struct __align__(32) LL32 {
  uint32_t w[8];
};

struct QrcpParams {
  double* B;
  int64_t ldb;
  int l;
  int64_t n;
  int kcap;
  int nb;           // effective block size = min(opts.nb, kcap)
  int nopivot;      // 1: plain (unpivoted) Householder QR -- the pivot of step s is the column at position s
  double atol, rtol;
  int cpc;          // columns per CTA
  int csm;          // of those, cached in shared memory
  int meta_smem;    // vn1/vn2/lpos in shared memory?
  double* vn1g;
  double* vn2g;
  int* lposg;
  LL16* rec;        // [2][G][l]        candidate columns
  LL32* inbox;      // [2][G dst][G src] headers
  uint32_t epoch;
  int64_t* jpvt;    // n, 1-based, LAPACK layout
  double* tau;      // kcap
  double* rdiag;    // kcap
  int* info;        // k, nsteps, nblocks, status, phase kilo-cycles...
  int* kbtrace;
  int kbcap;
  int* dbg;         // [G][8] per-CTA phase kilo-cycles (diagnostic)
  int lds;          // v2: column stride of the shared-memory slab (l rounded up to even: 16-byte aligned columns)
  int fast;         // v2: 1 = fixed 64-row chunk map with 128-bit accesses (l <= 576, aligned), 0 = generic loops
  long long* ts;    // v2 diagnostic: [G][16 warps][16 points] clock64 stamps of pivot step ts_step (BRA_QRCP_TS_STEP)
  int ts_step;
};

__device__ __forceinline__ void ll_store(LL16* p, uint32_t lo, uint32_t hi, uint32_t stamp) {
  asm volatile("st.relaxed.gpu.global.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(lo), "r"(stamp), "r"(hi),
               "r"(stamp)
               : "memory");
}
__device__ __forceinline__ void ll_store_d(LL16* p, double x, uint32_t stamp) {
  ll_store(p, (uint32_t)__double2loint(x), (uint32_t)__double2hiint(x), stamp);
}
__device__ __forceinline__ void ll_ld(const LL16* p, uint32_t& lo, uint32_t& s0, uint32_t& hi, uint32_t& s1) {
  asm volatile("ld.relaxed.gpu.global.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(lo), "=r"(s0), "=r"(hi), "=r"(s1)
               : "l"(p)
               : "memory");
}
__device__ __forceinline__ bool ll_load(const LL16* p, uint32_t stamp, uint32_t& lo, uint32_t& hi) {
  uint32_t s0, s1, spins = 0;
  do {
    ll_ld(p, lo, s0, hi, s1);
    if (s0 == stamp && s1 == stamp) return true;
  } while (++spins < SPIN_LIMIT);
  return false;
}

// header word: (v.lo, v.hi, lp, flag) as four 4-byte payloads, each followed by the step stamp
__device__ __forceinline__ void ll32_store(LL32* p, double v, int lp, int flag, uint32_t stamp) {
  asm volatile("st.relaxed.gpu.global.v8.b32 [%0], {%1,%2,%3,%2,%4,%2,%5,%2};" ::"l"(p),
               "r"((uint32_t)__double2loint(v)), "r"(stamp), "r"((uint32_t)__double2hiint(v)), "r"((uint32_t)lp),
               "r"((uint32_t)flag)
               : "memory");
}
// two doubles (rows r, r+1 of a record) as two adjacent LL16 words in one 32-byte store
__device__ __forceinline__ void ll32_store2(LL32* p, double x, double y, uint32_t stamp) {
  asm volatile("st.relaxed.gpu.global.v8.b32 [%0], {%1,%2,%3,%2,%4,%2,%5,%2};" ::"l"(p),
               "r"((uint32_t)__double2loint(x)), "r"(stamp), "r"((uint32_t)__double2hiint(x)),
               "r"((uint32_t)__double2loint(y)), "r"((uint32_t)__double2hiint(y))
               : "memory");
}
__device__ __forceinline__ void ll32_ld(const LL32* p, uint32_t (&q)[8]) {
  asm volatile("ld.relaxed.gpu.global.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(q[0]), "=r"(q[1]), "=r"(q[2]), "=r"(q[3]), "=r"(q[4]), "=r"(q[5]), "=r"(q[6]), "=r"(q[7])
               : "l"(p)
               : "memory");
}
__device__ __forceinline__ bool ll32_load(const LL32* p, uint32_t stamp, double& v, int& lp, int& flag) {
  uint32_t a, s0, b, s1, c, s2, d, s3, spins = 0;
  do {
    asm volatile("ld.relaxed.gpu.global.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(a), "=r"(s0), "=r"(b), "=r"(s1), "=r"(c), "=r"(s2), "=r"(d), "=r"(s3)
                 : "l"(p)
                 : "memory");
    if (s0 == stamp && s1 == stamp && s2 == stamp && s3 == stamp) {
      v = __hiloint2double((int)b, (int)a);
      lp = (int)c;
      flag = (int)d;
      return true;
    }
  } while (++spins < SPIN_LIMIT);
  return false;
}


}  // namespace braq
using namespace braq;

// qrcp_fast.cu
bool bra_qrcp_fast_plan(int l, int cpc, int nbe, size_t budget, bool aligned, int* jw, int* lds, int* csm, size_t* smem);
cudaError_t bra_qrcp_fast_launch(const QrcpParams& p, int G, int jw, size_t smem, cudaStream_t st);

// Diagnostic micro-benchmark (not part of libbrapprox): cost of handing a block of columns from one CTA to its ring
// neighbour, the per-step exchange of the one-sided Jacobi kernel (csrc/tail.cu).
//   path A (what the kernel does today): bulk store shared -> global, wait, st.release flag; the neighbour polls the
//           flag with ld.acquire and bulk-loads the block global -> shared (mbarrier complete_tx)
//   path B (candidate): inside a thread-block cluster the sender copies the block straight into the neighbour's shared
//           memory (cp.async.bulk.shared::cluster.shared::cta) and the copy engine completes the NEIGHBOUR's mbarrier;
//           a remote mbarrier arrive hands the slot back ("empty") to the sender
// CTAs 2i and 2i+1 swap a block every iteration (lock-step pairs, like one odd-even transposition step).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/dsmem_probe tools/dsmem_probe.cu && tools/dsmem_probe
// NOT YET RUN ON HARDWARE (written after the round's GPU minutes were spent): treat a hang or a wrong checksum as a bug
// in the probe first.  Run under `timeout 60`.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x)                                                                          \
  do {                                                                                 \
    cudaError_t e_ = (x);                                                              \
    if (e_ != cudaSuccess) {                                                           \
      std::printf("%s:%d %s -> %s\n", __FILE__, __LINE__, #x, cudaGetErrorString(e_)); \
      std::exit(1);                                                                    \
    }                                                                                  \
  } while (0)

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t n) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(b)), "r"(n) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* b, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive_local(uint64_t* b) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(b)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
  asm volatile(
      "{\n.reg .pred p;\nW: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D;\nbra W;\nD:\n}\n" ::"r"(s32(b)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ uint32_t mapa(uint32_t local, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ uint32_t cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t cluster_size() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// ---- path B: DSMEM exchange between the two CTAs of a pair (rank ^ 1) inside a cluster ---------------------------------
// full[2]: completed by the partner's copy into recv[b] (double buffered); empty: the partner arrives remotely when it
// has consumed my block, i.e. when `send` may be overwritten.
constexpr int MAXR = 32;      // doubles per thread held in registers (32 KB / 128 threads)
__global__ void __launch_bounds__(128, 1) k_dsmem(int iters, int bytes, long long* cyc, double* check) {
  extern __shared__ __align__(128) unsigned char sm[];
  double* send = reinterpret_cast<double*>(sm);
  double* recv0 = reinterpret_cast<double*>(sm + bytes);
  double* recv1 = reinterpret_cast<double*>(sm + 2 * (size_t)bytes);
  __shared__ __align__(8) uint64_t full[2], empty;
  const uint32_t rank = cluster_rank(), partner = rank ^ 1u;
  const int nd = bytes / 8;
  if (threadIdx.x == 0) {
    mbar_init(&full[0], 1);
    mbar_init(&full[1], 1);
    mbar_init(&empty, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = threadIdx.x; i < nd; i += blockDim.x) send[i] = (double)(blockIdx.x + 1);
  __syncthreads();
  cluster_sync();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    const int b = it & 1;
    double* rbuf = b ? recv1 : recv0;
    if (threadIdx.x == 0) {
      mbar_expect_tx(&full[b], (uint32_t)bytes);                     // my inbox b will receive `bytes`
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic writes to `send` before the async read
      const uint32_t dst = mapa(s32(rbuf), partner);                 // same offset in the partner's window
      const uint32_t bar = mapa(s32(&full[b]), partner);
      asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                   "r"(s32(send)), "r"((uint32_t)bytes), "r"(bar)
                   : "memory");
    }
    mbar_wait(&full[b], (it >> 1) & 1);                              // the partner's block has landed
    double v[MAXR];
    double acc = 0.0;
#pragma unroll
    for (int j = 0; j < MAXR; ++j) {
      const int i = threadIdx.x + j * 128;
      v[j] = (i < nd) ? rbuf[i] : 0.0;
      acc += v[j];
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      mbar_arrive_remote(mapa(s32(&empty), partner));                // I have consumed the partner's block
      mbar_wait(&empty, it & 1);                                     // the partner has consumed mine: `send` is free
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < MAXR; ++j) {
      const int i = threadIdx.x + j * 128;
      if (i < nd) send[i] = v[j] + 1.0;
    }
    __syncthreads();
    if (it == iters - 1 && threadIdx.x == 0) check[blockIdx.x] = acc;
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) cyc[blockIdx.x] = (t1 - t0) / iters;
  cluster_sync();                                                    // nobody exits while the partner may still write here
}

// ---- path B': PULL -- the consumer reads the partner's block out of the partner's shared memory with ld.shared::cluster
// after a `ready` arrive, and frees it with a `consumed` arrive (the protocol planned for the Jacobi kernel) -------------
__global__ void __launch_bounds__(128, 1) k_dsmem_pull(int iters, int bytes, long long* cyc, double* check) {
  extern __shared__ __align__(128) unsigned char sm[];
  double* send = reinterpret_cast<double*>(sm);
  __shared__ __align__(8) uint64_t ready, consumed;
  const uint32_t rank = cluster_rank(), partner = rank ^ 1u;
  const int nd = bytes / 8;
  if (threadIdx.x == 0) {
    mbar_init(&ready, 1);
    mbar_init(&consumed, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = threadIdx.x; i < nd; i += blockDim.x) send[i] = (double)(blockIdx.x + 1);
  __syncthreads();
  cluster_sync();
  const uint32_t rsend = mapa(s32(send), partner);
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    __syncthreads();                                                 // every thread's writes to `send` are done
    if (threadIdx.x == 0) {
      mbar_arrive_remote(mapa(s32(&ready), partner));                // release.cluster: my block may be read
      mbar_wait(&ready, it & 1);                                     // the partner's block may be read
    }
    __syncthreads();
    asm volatile("fence.acq_rel.cluster;" ::: "memory");
    double v[MAXR];
    double acc = 0.0;
#pragma unroll
    for (int j = 0; j < MAXR / 2; ++j) {
      const int i = 2 * (threadIdx.x + j * 128);                     // 16-byte remote loads
      double a = 0.0, c = 0.0;
      if (i < nd) asm volatile("ld.shared::cluster.v2.f64 {%0, %1}, [%2];" : "=d"(a), "=d"(c) : "r"(rsend + 8u * (uint32_t)i) : "memory");
      v[2 * j] = a;
      v[2 * j + 1] = c;
      acc += a + c;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      mbar_arrive_remote(mapa(s32(&consumed), partner));             // the partner may overwrite its block
      mbar_wait(&consumed, it & 1);                                  // I may overwrite mine
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < MAXR / 2; ++j) {
      const int i = 2 * (threadIdx.x + j * 128);
      if (i < nd) {
        send[i] = v[2 * j] + 1.0;
        send[i + 1] = v[2 * j + 1] + 1.0;
      }
    }
    if (it == iters - 1 && threadIdx.x == 0) check[blockIdx.x] = acc;
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) cyc[blockIdx.x] = (t1 - t0) / iters;
  cluster_sync();
}

// ---- path A: the same pair exchange through global memory (bulk store + flag, poll + bulk load) ---------------------
__global__ void __launch_bounds__(128, 1) k_global(int iters, int bytes, double* gbuf, unsigned* flags, long long* cyc,
                                                   double* check) {
  extern __shared__ __align__(128) unsigned char sm[];
  double* send = reinterpret_cast<double*>(sm);
  double* rbuf = reinterpret_cast<double*>(sm + bytes);
  __shared__ __align__(8) uint64_t full;
  const int me = blockIdx.x, partner = me ^ 1;
  const int nd = bytes / 8;
  if (threadIdx.x == 0) {
    mbar_init(&full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = threadIdx.x; i < nd; i += blockDim.x) send[i] = (double)(me + 1);
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    // two global buffers per CTA, alternating: the partner has finished loading buffer (it & 1) of iteration it - 2
    // before it raised the flag this CTA waited for in iteration it - 1
    double* mine = gbuf + ((size_t)me * 2 + (it & 1)) * nd;
    const double* theirs = gbuf + ((size_t)partner * 2 + (it & 1)) * nd;
    if (threadIdx.x == 0) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(mine), "r"(s32(send)),
                   "r"((uint32_t)bytes)
                   : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
      asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(flags + me), "r"((unsigned)(it + 1)) : "memory");
      unsigned v;
      do {
        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(flags + partner) : "memory");
      } while (v < (unsigned)(it + 1));
      mbar_expect_tx(&full, (uint32_t)bytes);
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s32(rbuf)),
                   "l"(theirs), "r"((uint32_t)bytes), "r"(s32(&full))
                   : "memory");
    }
    mbar_wait(&full, it & 1);
    double acc = 0.0;
    for (int i = threadIdx.x; i < nd; i += blockDim.x) acc += rbuf[i];
    __syncthreads();
    for (int i = threadIdx.x; i < nd; i += blockDim.x) send[i] = rbuf[i] + 1.0;
    __syncthreads();
    if (it == iters - 1 && threadIdx.x == 0) check[me] = acc;
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) cyc[me] = (t1 - t0) / iters;
}

int main() {
  const int iters = 2000;
  long long* cyc;
  double* check;
  CK(cudaMallocManaged(&cyc, 256 * sizeof(long long)));
  CK(cudaMallocManaged(&check, 256 * sizeof(double)));
  for (int bytes : {4096, 16384, 32768}) {
    // path A, 64 CTAs (the Jacobi grid at k = 500), all co-resident (cooperative launch not needed: 64 <= 148, 1 CTA/SM)
    {
      const int G = 64;
      double* gbuf;
      unsigned* flags;
      CK(cudaMalloc(&gbuf, (size_t)G * 2 * bytes));
      CK(cudaMalloc(&flags, G * sizeof(unsigned)));
      CK(cudaMemset(flags, 0, G * sizeof(unsigned)));
      CK(cudaFuncSetAttribute(k_global, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * bytes));
      k_global<<<G, 128, 2 * bytes>>>(iters, bytes, gbuf, flags, cyc, check);
      CK(cudaDeviceSynchronize());
      long long mx = 0;
      for (int i = 0; i < G; ++i) mx = cyc[i] > mx ? cyc[i] : mx;
      std::printf("global pair  %6d B: %lld cycles / exchange (max over %d CTAs), check %.0f\n", bytes, mx, G, check[0]);
      CK(cudaFree(gbuf));
      CK(cudaFree(flags));
    }
    for (int cs : {2, 4, 8}) {
      const int G = 64;
      CK(cudaFuncSetAttribute(k_dsmem, cudaFuncAttributeMaxDynamicSharedMemorySize, 3 * bytes));
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(G);
      cfg.blockDim = dim3(128);
      cfg.dynamicSmemBytes = 3 * (size_t)bytes;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension;
      at[0].val.clusterDim.x = cs;
      at[0].val.clusterDim.y = 1;
      at[0].val.clusterDim.z = 1;
      cfg.attrs = at;
      cfg.numAttrs = 1;
      CK(cudaLaunchKernelEx(&cfg, k_dsmem, iters, bytes, cyc, check));
      CK(cudaDeviceSynchronize());
      long long mx = 0;
      for (int i = 0; i < G; ++i) mx = cyc[i] > mx ? cyc[i] : mx;
      std::printf("dsmem push   %6d B, cluster %d: %lld cycles / exchange (max over %d CTAs), check %.0f\n", bytes, cs, mx, G,
                  check[0]);
      cfg.dynamicSmemBytes = (size_t)bytes;
      CK(cudaFuncSetAttribute(k_dsmem_pull, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
      CK(cudaLaunchKernelEx(&cfg, k_dsmem_pull, iters, bytes, cyc, check));
      CK(cudaDeviceSynchronize());
      mx = 0;
      for (int i = 0; i < G; ++i) mx = cyc[i] > mx ? cyc[i] : mx;
      std::printf("dsmem pull   %6d B, cluster %d: %lld cycles / exchange (max over %d CTAs), check %.0f\n", bytes, cs, mx, G,
                  check[0]);
    }
  }
  return 0;
}

// mc_probe.cu -- measurement probe (not part of the product): does TMA multicast relieve the L2->SM operand path?
//
// Every CTA streams 16 KB "weight" boxes that ALL CTAs read (L2-resident, like the B operand of the fused WN layer) through a
// 4-stage ring, optionally next to 16 KB of per-CTA "activation" boxes (like the A operand).  Modes:
//   cluster 1: every CTA loads its own copy (unicast)
//   cluster C: every CTA loads 1/C of the box and multicasts it to the C CTAs of its cluster
// Also prints how many clusters of each size are co-resident with the layer kernel's footprint (320 threads, ~200 KB smem).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 tools/mc_probe.cu -o /tmp/mc_probe -lcuda && /tmp/mc_probe
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s -> %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

__device__ __forceinline__ uint32_t s32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(b)), "r"(c) : "memory"); }
__device__ __forceinline__ void mbar_expect(uint64_t* b, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(b)), "r"(bytes) : "memory"); }
__device__ __forceinline__ bool mbar_try(uint64_t* b, uint32_t par) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(s32(b)), "r"(par) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t par) {
  const long long t0 = clock64();
  while (!mbar_try(b, par))
    if (clock64() - t0 > 4000000000LL) { printf("probe: barrier timeout block %d\n", blockIdx.x); __trap(); }
}
__device__ __forceinline__ uint32_t cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void arrive_remote(uint64_t* b, uint32_t rank) {
  uint32_t a;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(a) : "r"(s32(b)), "r"(rank));
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(a) : "memory");
}
__device__ __forceinline__ void tma2d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(s32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(s32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma2d_mc(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, uint16_t mask) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;"
               ::"r"(s32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(s32(bar)), "r"(c0), "r"(c1), "h"(mask) : "memory");
}

constexpr int kStages = 4;
constexpr int kBox = 16384;

// tm_w: {768, 512} 16-bit, box {64, 128 / C}; tm_a: {64, rows} 16-bit, box {64, 128}
template <int C>
__global__ void __launch_bounds__(128, 1) probe_kernel(const __grid_constant__ CUtensorMap tm_w, const __grid_constant__ CUtensorMap tm_a,
                                                       int iters, int with_a, int a_rows_per_cta) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + kStages * 2 * kBox);
  uint64_t* empty = full + kStages;
  const uint32_t rank = C > 1 ? cluster_rank() : 0;
  if (threadIdx.x == 0) {
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], C);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (C > 1) cluster_sync(); else __syncthreads();
  if (threadIdx.x == 0) {
    // producer + consumer in one thread, software-pipelined: issue stage i, then retire stage i - (kStages - 1)
    for (int i = 0; i < iters + kStages - 1; ++i) {
      if (i < iters) {
        const int s = i % kStages;
        if (i >= kStages) mbar_wait(&empty[s], ((i / kStages) & 1) ^ 1);
        uint8_t* dst = smem + s * 2 * kBox;
        mbar_expect(&full[s], kBox + (with_a ? kBox : 0));
        const int j = i % 12, rb = (i / 12) % 4;
        if (C == 1) {
          tma2d(dst, &tm_w, &full[s], j * 64, rb * 128);
        } else {
          constexpr int rows = 128 / C;
          tma2d_mc(dst + rank * rows * 128, &tm_w, &full[s], j * 64, rb * 128 + static_cast<int>(rank) * rows, static_cast<uint16_t>((1u << C) - 1));
        }
        if (with_a) tma2d(dst + kBox, &tm_a, &full[s], 0, blockIdx.x * a_rows_per_cta + (i * 128) % a_rows_per_cta);
      }
      const int r = i - (kStages - 1);
      if (r >= 0) {
        const int s = r % kStages;
        mbar_wait(&full[s], (r / kStages) & 1);
        if (C == 1) {
          asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(&empty[s])) : "memory");
        } else {
          for (uint32_t c = 0; c < C; ++c) arrive_remote(&empty[s], c);   // every CTA that writes into my stage s may go on
        }
      }
    }
  }
  if (C > 1) cluster_sync(); else __syncthreads();
}

typedef CUresult (*PFN_encode)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                               const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static CUtensorMap make2d(PFN_encode enc, void* base, uint64_t d0, uint64_t d1, uint32_t b0, uint32_t b1) {
  CUtensorMap m;
  cuuint64_t gd[2] = {d0, d1}, gs[1] = {d0 * 2};
  cuuint32_t bx[2] = {b0, b1}, es[2] = {1, 1};
  CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, base, gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); exit(1); }
  return m;
}

template <int C>
static void run(PFN_encode enc, void* d_w, void* d_a, int n_sms, int iters, int with_a, int a_rows_per_cta) {
  CUtensorMap tm_w = make2d(enc, d_w, 768, 512, 64, 128 / C);
  CUtensorMap tm_a = make2d(enc, d_a, 64, static_cast<uint64_t>(n_sms) * a_rows_per_cta, 64, 128);
  const int smem = kStages * 2 * kBox + 1024 + 256;
  CK(cudaFuncSetAttribute(probe_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(n_sms / C * C);
  cfg.blockDim = dim3(128);
  cfg.dynamicSmemBytes = 200 * 1024;   // one CTA per SM, like the layer kernel
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = C;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  (void)smem;
  int max_clusters = 0;
  CK(cudaOccupancyMaxActiveClusters(&max_clusters, probe_kernel<C>, &cfg));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  float best = 1e30f;
  for (int rep = 0; rep < 5; ++rep) {
    CK(cudaEventRecord(e0));
    CK(cudaLaunchKernelEx(&cfg, probe_kernel<C>, tm_w, tm_a, iters, with_a, a_rows_per_cta));
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    if (rep > 0 && ms < best) best = ms;
  }
  const double n_cta = n_sms / C * C;
  const double bytes = n_cta * iters * (kBox + (with_a ? kBox : 0));
  printf("cluster %d  with_a %d: max co-resident clusters %d (= %d SMs) | %.3f ms  smem fill %.2f TB/s = %.1f B/ns/SM | L2 reads if multicast is one read: %.2f TB/s\n",
         C, with_a, max_clusters, max_clusters * C, best, bytes / best * 1e-9, bytes / best * 1e-6 / n_cta,
         n_cta * iters * (kBox / C + (with_a ? kBox : 0)) / best * 1e-9);
}

int main() {
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  printf("%s, %d SMs\n", prop.name, prop.multiProcessorCount);
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
  PFN_encode enc = reinterpret_cast<PFN_encode>(fn);
  const int n_sms = prop.multiProcessorCount;
  const int a_rows_per_cta = 4096;   // 512 KB per CTA -> 74 MB: L2-resident
  void *d_w, *d_a;
  CK(cudaMalloc(&d_w, 768 * 512 * 2));
  CK(cudaMalloc(&d_a, static_cast<size_t>(n_sms) * a_rows_per_cta * 128));
  CK(cudaMemset(d_w, 0, 768 * 512 * 2));
  CK(cudaMemset(d_a, 0, static_cast<size_t>(n_sms) * a_rows_per_cta * 128));
  const int iters = 20000;
  for (int with_a = 0; with_a < 2; ++with_a) {
    run<1>(enc, d_w, d_a, n_sms, iters, with_a, a_rows_per_cta);
    run<2>(enc, d_w, d_a, n_sms, iters, with_a, a_rows_per_cta);
    run<4>(enc, d_w, d_a, n_sms, iters, with_a, a_rows_per_cta);
  }
  return 0;
}

// mma_pair_bench.cu - cycles per tcgen05.mma.cta_group::2 (M = 256 across a CTA pair, K = 16, bf16) as a function of N,
// issued back to back by the leader CTA; both CTAs hold their 128 rows of A and N/2 rows of B in shared memory.
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t desc_sw128(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3fffu);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

__global__ void __launch_bounds__(128, 1) mma_pair_bench(int N, int n_mma, unsigned long long* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~(uintptr_t)1023);
  uint32_t* slot = reinterpret_cast<uint32_t*>(smem + 16384 + 32768);
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 16384 + 32768 + 64);
  const int warp = threadIdx.x >> 5;
  uint32_t rank;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::);
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *slot;
  if (threadIdx.x == 0 && rank == 0) {
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
    const uint64_t ad = desc_sw128(smem_u32(smem)), bd = desc_sw128(smem_u32(smem + 16384));
    const long long t0 = clock64();
    for (int i = 0; i < n_mma; ++i) {
      const uint32_t acc = i != 0;
      asm volatile(
          "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem),
          "l"(ad + 2 * (i & 3)), "l"(bd + 2 * (i & 3)), "r"(idesc), "r"(acc)
          : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
                 "h"((uint16_t)3)
                 : "memory");
    const long long t1 = clock64();
    uint32_t ok = 0;
    while (!ok) {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(bar)) : "memory");
    }
    const long long t2 = clock64();
    out[0] = (unsigned long long)(t2 - t0);
    out[1] = (unsigned long long)(t1 - t0);
  }
  if (threadIdx.x == 0 && rank == 1) {   // the peer waits for the multicast commit too, so it does not exit early
    uint32_t ok = 0;
    while (!ok) {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(bar)) : "memory");
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}

int main() {
  unsigned long long* out;
  cudaMalloc(&out, 64);
  unsigned long long h[2];
  cudaFuncSetAttribute(mma_pair_bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  for (int N : {32, 64, 96, 128, 192, 256}) {
    const int n_mma = 512;
    for (int rep = 0; rep < 2; ++rep) {
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(2);
      cfg.blockDim = dim3(128);
      cfg.dynamicSmemBytes = 64 * 1024;
      cudaLaunchAttribute attr[1];
      attr[0].id = cudaLaunchAttributeClusterDimension;
      attr[0].val.clusterDim.x = 2;
      attr[0].val.clusterDim.y = 1;
      attr[0].val.clusterDim.z = 1;
      cfg.attrs = attr;
      cfg.numAttrs = 1;
      cudaLaunchKernelEx(&cfg, mma_pair_bench, N, n_mma, out);
    }
    cudaDeviceSynchronize();
    cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost);
    printf("tcgen05.mma cta_group::2 M=256 N=%3d K=16: %.1f cycles per MMA (nominal %.0f per SM), issue loop %.1f\n", N,
           (double)h[0] / n_mma, 128.0 * N * 16 / 4096.0, (double)h[1] / n_mma);
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}

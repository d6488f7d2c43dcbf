// tmem_mma_bench.cu - two numbers the epilogue / fusion designs depend on, measured on one SM of a B200:
//   (1) tcgen05.ld throughput (bytes per cycle per SM) for 4 / 8 / 16 reading warps and x32 / x64 / x128 shapes
//   (2) cycles per tcgen05.mma (cta_group::1, M = 128, K = 16, bf16) as a function of N, issued back to back
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_mma_bench tmem_mma_bench.cu ; run on a B200.
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

#define LD_X32(taddr)                                                                                                    \
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];" \
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31]) \
               : "r"(taddr))
#define LD_X16(taddr)                                                                                                    \
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];" \
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]) \
               : "r"(taddr))

// mode 0: x32 + wait each; mode 1: 2 x x32 then wait; mode 2: x16 + wait each
__global__ void __launch_bounds__(512, 1) ld_bench(int nwarps, int reps, int mode, unsigned long long* out, uint32_t* sink) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = slot + ((uint32_t)((warp & 3) * 32) << 16);
  uint32_t acc = 0;
  __syncthreads();
  const long long t0 = clock64();
  if (warp < nwarps) {
    uint32_t r[32];
    uint32_t r2[32];
    for (int i = 0; i < reps; ++i) {
      const uint32_t col = (uint32_t)((i * 64 + (warp >> 2) * 32) & 511);
      if (mode == 0) {
        LD_X32(base + col);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        acc ^= r[0] ^ r[31];
      } else if (mode == 1) {
        LD_X32(base + col);
        {
          uint32_t* rr = r2;
          asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                       : "=r"(rr[0]), "=r"(rr[1]), "=r"(rr[2]), "=r"(rr[3]), "=r"(rr[4]), "=r"(rr[5]), "=r"(rr[6]), "=r"(rr[7]), "=r"(rr[8]), "=r"(rr[9]), "=r"(rr[10]), "=r"(rr[11]), "=r"(rr[12]), "=r"(rr[13]), "=r"(rr[14]), "=r"(rr[15]), "=r"(rr[16]), "=r"(rr[17]), "=r"(rr[18]), "=r"(rr[19]), "=r"(rr[20]), "=r"(rr[21]), "=r"(rr[22]), "=r"(rr[23]), "=r"(rr[24]), "=r"(rr[25]), "=r"(rr[26]), "=r"(rr[27]), "=r"(rr[28]), "=r"(rr[29]), "=r"(rr[30]), "=r"(rr[31])
                       : "r"(base + ((col + 32) & 511)));
        }
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        acc ^= r[0] ^ r2[31];
      } else {
        LD_X16(base + col);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        acc ^= r[0] ^ r[15];
      }
    }
  }
  __syncthreads();
  const long long t1 = clock64();
  if (threadIdx.x == 0) out[0] = (unsigned long long)(t1 - t0);
  if (acc == 0x12345678u) sink[threadIdx.x] = acc;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(slot), "r"(512));
}

__device__ __forceinline__ uint64_t desc_sw128(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3fffu);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

// one thread issues `n_mma` MMAs (M=128, N, K=16) back to back, then commits; cycles until the commit lands
__global__ void __launch_bounds__(128, 1) mma_bench(int N, int n_mma, unsigned long long* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~(uintptr_t)1023);
  __shared__ uint32_t slot;
  __shared__ uint64_t bar;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (threadIdx.x == 0) {
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint64_t ad = desc_sw128(smem_u32(smem)), bd = desc_sw128(smem_u32(smem + 16384));
    const long long t0 = clock64();
    for (int i = 0; i < n_mma; ++i) {
      const uint32_t acc = i != 0;
      asm volatile(
          "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(slot),
          "l"(ad + 2 * (i & 3)), "l"(bd + 2 * (i & 3)), "r"(idesc), "r"(acc)
          : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    const long long t1 = clock64();
    uint32_t ok = 0;
    while (!ok) {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(&bar)) : "memory");
    }
    const long long t2 = clock64();
    out[0] = (unsigned long long)(t2 - t0);
    out[1] = (unsigned long long)(t1 - t0);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(slot), "r"(512));
}

int main() {
  unsigned long long* out;
  uint32_t* sink;
  cudaMalloc(&out, 64);
  cudaMalloc(&sink, 4096);
  unsigned long long h[2];
  const int reps = 2000;
  const char* names[3] = {"x32+wait", "2*x32+wait", "x16+wait"};
  for (int mode = 0; mode < 3; ++mode)
    for (int nw : {4, 8, 16}) {
      for (int rep = 0; rep < 2; ++rep) ld_bench<<<1, 512>>>(nw, reps, mode, out, sink);
      cudaDeviceSynchronize();
      cudaMemcpy(h, out, 8, cudaMemcpyDeviceToHost);
      const double bytes = (double)nw * reps * (mode == 1 ? 8192 : (mode == 2 ? 2048 : 4096));
      printf("tcgen05.ld %-11s warps=%2d  %8llu cycles  %.1f B/cycle/SM  (%.0f cycles per warp-load)\n", names[mode], nw, h[0],
             bytes / (double)h[0], (double)h[0] / reps);
    }
  cudaFuncSetAttribute(mma_bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  for (int N : {32, 64, 96, 128, 192, 256}) {
    const int n_mma = 512;
    for (int rep = 0; rep < 2; ++rep) mma_bench<<<1, 128, 64 * 1024>>>(N, n_mma, out);
    cudaDeviceSynchronize();
    cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost);
    printf("tcgen05.mma cta_group::1 M=128 N=%3d K=16: %.1f cycles per MMA (nominal %.0f), issue loop %.1f cycles per MMA\n", N,
           (double)h[0] / n_mma, 128.0 * N * 16 / 4096.0, (double)h[1] / n_mma);
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}

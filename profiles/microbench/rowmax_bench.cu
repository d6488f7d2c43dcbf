// rowmax_bench.cu - (1) the register <-> (lane, column) mapping of tcgen05.ld.16x256b, read back from a TMEM tile filled with
// tcgen05.st.32x32b (value = lane * 1000 + column), and (2) the cost of the 32-row max of a 32x32 accumulator block for
//   mode 0: tcgen05.ld.32x32b.x32 + 32 x redux.sync.max.f32                      (round-1 epilogues)
//   mode 1: 2 x tcgen05.ld.16x256b.x4 + 24 thread-local FMNMX + 7-shuffle halving butterfly (rows r, r+8, r+16, r+24 of
//           a column pair already sit in ONE thread, so only 8 row classes are reduced across lanes)
//   mode 2: tcgen05.ld.32x32b.x32 + 31-shuffle transpose
// with 4 / 8 warps on one SM.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o rowmax_bench rowmax_bench.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
               : "r"(taddr));
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}
// 16 lanes x 32 columns: 16 registers per thread
__device__ __forceinline__ void ld16x256_x4(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile("tcgen05.ld.sync.aligned.16x256b.x4.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
               : "r"(taddr));
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void st32(uint32_t taddr, const float (&v)[32]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};" ::"r"(taddr),
               "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])), "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])),
               "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])), "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])), "r"(__float_as_uint(v[15])),
               "r"(__float_as_uint(v[16])), "r"(__float_as_uint(v[17])), "r"(__float_as_uint(v[18])), "r"(__float_as_uint(v[19])), "r"(__float_as_uint(v[20])), "r"(__float_as_uint(v[21])), "r"(__float_as_uint(v[22])), "r"(__float_as_uint(v[23])),
               "r"(__float_as_uint(v[24])), "r"(__float_as_uint(v[25])), "r"(__float_as_uint(v[26])), "r"(__float_as_uint(v[27])), "r"(__float_as_uint(v[28])), "r"(__float_as_uint(v[29])), "r"(__float_as_uint(v[30])), "r"(__float_as_uint(v[31]))
               : "memory");
}
__device__ __forceinline__ void wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ float rows_max_redux(float (&v)[32], int lane) {
  float m = 0.f;
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    float r;
    asm volatile("redux.sync.max.f32 %0, %1, 0xffffffff;" : "=f"(r) : "f"(v[j]));
    if (lane == j) m = r;
  }
  return m;
}
__device__ __forceinline__ float rows_max_shfl(float (&v)[32], int lane) {
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    const bool hi = (lane & off) != 0;
#pragma unroll
    for (int j = 0; j < off; ++j) {
      const float send = hi ? v[j] : v[j + off];
      const float keep = hi ? v[j + off] : v[j];
      v[j] = fmaxf(keep, __shfl_xor_sync(0xffffffffu, send, off));
    }
  }
  return v[0];
}
// Max over the 32 lanes of the warp's TMEM quarter of the 32 columns starting at `taddr` (lane field = the quarter's first lane).
// Assumed 16x256b.xN fragment: register 4s+{0,1} = (lane t/4,     columns 8s + 2(t%4) + {0,1}),
//                              register 4s+{2,3} = (lane t/4 + 8, same columns)               - checked by the dump below.
// Returns in lane t the max of column  col_of_lane(t) = 8*((t>>3)&3) ... see the host-side check; the caller writes
// out[col_of_lane].  7 shuffles.
__device__ __forceinline__ float rows_max_16x256(uint32_t taddr, int lane, int& col) {
  float a[16], b[16];
  ld16x256_x4(taddr, a);                              // lanes +0..15
  ld16x256_x4(taddr + (16u << 16), b);                // lanes +16..31
  wait_ld();
  float m[8];                                         // m[2s+e]: column 8s + 2(t%4) + e, max over rows {t/4, +8, +16, +24}
#pragma unroll
  for (int s = 0; s < 4; ++s) {
#pragma unroll
    for (int e = 0; e < 2; ++e) m[2 * s + e] = fmaxf(fmaxf(a[4 * s + e], a[4 * s + 2 + e]), fmaxf(b[4 * s + e], b[4 * s + 2 + e]));
  }
  // halving butterfly over the 8 row classes (lane bits 2, 3, 4): each step exchanges half of the remaining values
  {
    const bool hi = (lane & 16) != 0;                 // keep column sub-blocks {2,3} if hi else {0,1}
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float send = hi ? m[j] : m[j + 4];
      const float keep = hi ? m[j + 4] : m[j];
      m[j] = fmaxf(keep, __shfl_xor_sync(0xffffffffu, send, 16));
    }
  }
  {
    const bool hi = (lane & 8) != 0;                  // of the two remaining sub-blocks keep the second if hi
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const float send = hi ? m[j] : m[j + 2];
      const float keep = hi ? m[j + 2] : m[j];
      m[j] = fmaxf(keep, __shfl_xor_sync(0xffffffffu, send, 8));
    }
  }
  {
    const bool hi = (lane & 4) != 0;                  // of the column pair keep the odd one if hi
    const float send = hi ? m[0] : m[1];
    const float keep = hi ? m[1] : m[0];
    m[0] = fmaxf(keep, __shfl_xor_sync(0xffffffffu, send, 4));
  }
  col = 8 * (2 * ((lane >> 4) & 1) + ((lane >> 3) & 1)) + 2 * (lane & 3) + ((lane >> 2) & 1);
  return m[0];
}

__global__ void __launch_bounds__(512, 1) bench(int mode, int nwarps, int reps, unsigned long long* cyc, float* out, int* dump) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t q = (uint32_t)(warp & 3);
  const uint32_t base = slot + ((q * 32u) << 16);
  // fill: TMEM[lane L (global), column c] = L * 1000 + c  for all 512 columns (warps 0..3, one lane quarter each)
  if (warp < 4) {
    for (int c0 = 0; c0 < 512; c0 += 32) {
      float v[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = (float)((int)(q * 32 + lane) * 1000 + c0 + j);
      st32(base + c0, v);
    }
    wait_st();
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (dump && warp == 1) {     // lane quarter 1 (global lanes 32..63), columns 64..95: what does each register hold?
    float a[16], b[16];
    ld16x256_x4(base + 64, a);
    ld16x256_x4(base + 64 + (16u << 16), b);
    wait_ld();
    for (int i = 0; i < 16; ++i) {
      dump[(lane * 32 + i) * 2 + 0] = (int)a[i] / 1000;
      dump[(lane * 32 + i) * 2 + 1] = (int)a[i] % 1000;
      dump[(lane * 32 + 16 + i) * 2 + 0] = (int)b[i] / 1000;
      dump[(lane * 32 + 16 + i) * 2 + 1] = (int)b[i] % 1000;
    }
    int col;
    const float m = rows_max_16x256(base + 64, lane, col);
    dump[32 * 32 * 2 + lane * 2] = col;
    dump[32 * 32 * 2 + lane * 2 + 1] = (int)m;      // expected: (32 + 31) * 1000 + 64 + col
  }
  float acc = 0.f;
  __syncthreads();
  const long long t0 = clock64();
  if (warp < nwarps) {
    for (int i = 0; i < reps; ++i) {
      const uint32_t col0 = (uint32_t)((i * 64 + (warp >> 2) * 32) & 511);
      if (mode == 0) {
        float v[32];
        ld32(base + col0, v);
        wait_ld();
        acc += rows_max_redux(v, lane);
      } else if (mode == 1) {
        int col;
        acc += rows_max_16x256(base + col0, lane, col);
      } else {
        float v[32];
        ld32(base + col0, v);
        wait_ld();
        acc += rows_max_shfl(v, lane);
      }
    }
  }
  __syncthreads();
  const long long t1 = clock64();
  if (threadIdx.x == 0) cyc[0] = (unsigned long long)(t1 - t0);
  out[threadIdx.x] = acc;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(slot), "r"(512));
}

int main() {
  unsigned long long* cyc;
  float* out;
  int* dump;
  cudaMalloc(&cyc, 8);
  cudaMalloc(&out, 512 * 4);
  cudaMalloc(&dump, (32 * 32 * 2 + 64) * 4);
  cudaMemset(dump, 0, (32 * 32 * 2 + 64) * 4);
  bench<<<1, 512>>>(1, 8, 1, cyc, out, dump);
  cudaDeviceSynchronize();
  static int h[32 * 32 * 2 + 64];
  cudaMemcpy(h, dump, sizeof(h), cudaMemcpyDeviceToHost);
  printf("tcgen05.ld.16x256b.x4 fragment (warp 1: TMEM lanes 32..63, columns 64..95); entries are (lane - 32, column - 64)\n");
  int bad = 0;
  for (int t = 0; t < 32; ++t) {
    if (t < 8 || t == 31) printf("thread %2d:", t);
    for (int half = 0; half < 2; ++half)
      for (int i = 0; i < 16; ++i) {
        const int L = h[(t * 32 + half * 16 + i) * 2] - 32, C = h[(t * 32 + half * 16 + i) * 2 + 1] - 64;
        const int s = i / 4, e = i & 1, up = (i >> 1) & 1;
        const int eL = 16 * half + t / 4 + 8 * up, eC = 8 * s + 2 * (t & 3) + e;
        if (L != eL || C != eC) ++bad;
        if (t < 8 || t == 31) printf(" r%d=(%d,%d)", half * 16 + i, L, C);
      }
    if (t < 8 || t == 31) printf("\n");
  }
  printf("fragment matches the assumed m16n8-style layout: %s (%d mismatches)\n", bad ? "NO" : "yes", bad);
  int badm = 0;
  for (int t = 0; t < 32; ++t) {
    const int col = h[32 * 32 * 2 + t * 2], m = h[32 * 32 * 2 + t * 2 + 1];
    if (m != 63 * 1000 + 64 + col) ++badm;
  }
  printf("rows_max_16x256: every lane holds the max of its column: %s (%d wrong); lane->column:", badm ? "NO" : "yes", badm);
  for (int t = 0; t < 32; ++t) printf(" %d", h[32 * 32 * 2 + t * 2]);
  printf("\n");
  const char* names[3] = {"32x32b + 32 redux.sync.max", "2 x 16x256b + local max + 7 shfl", "32x32b + 31 shfl transpose"};
  for (int mode = 0; mode < 3; ++mode)
    for (int nw : {1, 4, 8, 16}) {
      const int reps = 500;
      for (int r = 0; r < 2; ++r) bench<<<1, 512>>>(mode, nw, reps, cyc, out, nullptr);
      cudaDeviceSynchronize();
      unsigned long long hc;
      cudaMemcpy(&hc, cyc, 8, cudaMemcpyDeviceToHost);
      printf("%-36s warps=%2d: %.0f cycles per 32x32 block per warp, %.1f cycles per block per SM (TMEM load included)\n", names[mode], nw,
             (double)hc / reps, (double)hc / reps / nw);
    }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}

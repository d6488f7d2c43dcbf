// redux_bench.cu - throughput of redux.sync.max.f32 (SASS CREDUX.MAX.F32) vs a 31-shuffle lane-transpose max,
// per 32x32 block (32 columns reduced over the warp's 32 lanes), for 4/8/16 warps on one SM.
#include <cstdio>
#include <cuda_runtime.h>

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
// smem transpose: lane writes its row (pitch 33 words), then reads column `lane`
__device__ __forceinline__ float rows_max_smem(float (&v)[32], int lane, float* buf) {
#pragma unroll
  for (int j = 0; j < 32; ++j) buf[lane * 33 + j] = v[j];
  __syncwarp();
  float m = buf[lane];
#pragma unroll
  for (int r = 1; r < 32; ++r) m = fmaxf(m, buf[r * 33 + lane]);
  __syncwarp();
  return m;
}

__global__ void __launch_bounds__(512, 1) bench(int mode, int nwarps, int reps, const float* in, float* out, unsigned long long* cyc) {
  __shared__ float buf[8][32 * 33];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float v[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) v[j] = in[threadIdx.x * 32 + j];
  float acc = 0.f;
  __syncthreads();
  const long long t0 = clock64();
  if (warp < nwarps) {
    for (int i = 0; i < reps; ++i) {
      float w[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) w[j] = v[j] + acc;
      acc += mode == 0 ? rows_max_redux(w, lane) : (mode == 1 ? rows_max_shfl(w, lane) : rows_max_smem(w, lane, buf[warp & 7]));
    }
  }
  __syncthreads();
  const long long t1 = clock64();
  if (threadIdx.x == 0) cyc[0] = (unsigned long long)(t1 - t0);
  out[threadIdx.x] = acc;
}

int main() {
  float *in, *out;
  unsigned long long* cyc;
  cudaMalloc(&in, 512 * 32 * 4);
  cudaMemset(in, 0, 512 * 32 * 4);
  cudaMalloc(&out, 512 * 4);
  cudaMalloc(&cyc, 8);
  const char* names[3] = {"redux.sync.max.f32", "shfl transpose", "smem transpose"};
  for (int mode = 0; mode < 3; ++mode)
    for (int nw : {1, 4, 8}) {
      const int reps = 500;
      for (int r = 0; r < 2; ++r) bench<<<1, 512>>>(mode, nw, reps, in, out, cyc);
      cudaDeviceSynchronize();
      unsigned long long h;
      cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
      printf("%-20s warps=%2d: %.0f cycles per 32x32 block per warp, %.0f cycles per block per SM\n", names[mode], nw, (double)h / reps,
             (double)h / reps / nw);
    }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
}

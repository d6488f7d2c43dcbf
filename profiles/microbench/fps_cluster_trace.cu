// fps_cluster_trace.cu - where an FPS iteration of a CLUSTER-per-cloud launch (N > 8192, BASELINE configs[3]) spends its time.
// Includes the product kernel (csrc/fps.cu) with P3TOK_FPS_TRACE: iteration 100 of cloud 0 stamps clock64() in warp 0 and
// warp 5 of every CTA of the cluster at: 0 loop top, 1 distances + warp argmax done, 2 all warps' candidates in (warp 0) /
// barrier passed, 3 candidate pushed to the peers, 4 all peers' candidates received, 5 winner published, 6 winner read.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -DP3TOK_FPS_TRACE -I ../../include
//        -I ../../adapting-2d-vits-for-3d-point-cloud-understanding_b200/csrc -o _bin/fps_cluster_trace fps_cluster_trace.cu
// Run:   [P3TOK_FPS_CLUSTER=n] ./_bin/fps_cluster_trace [N] [G] [B]      (n forces the cluster size; default: chosen by p3tok_fps)
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "fps.cu"

namespace p3tok {
void set_error(const char* fmt, ...) { va_list a; va_start(a, fmt); vfprintf(stderr, fmt, a); va_end(a); fprintf(stderr, "\n"); }
int cuda_fail(cudaError_t e, const char* what) { fprintf(stderr, "CUDA error %s at %s\n", cudaGetErrorString(e), what); return P3TOK_ERR_CUDA; }
void count_launch(int) {}
}  // namespace p3tok

int main(int argc, char** argv) {
  const int N = argc > 1 ? atoi(argv[1]) : 65536, G = argc > 2 ? atoi(argv[2]) : 2048, B = argc > 3 ? atoi(argv[3]) : 16;
  std::vector<float> h((size_t)B * N * 3);
  unsigned s = 12345u;
  for (auto& v : h) { s = s * 1664525u + 1013904223u; v = (float)(s >> 8) / 8388608.f - 1.f; }
  std::vector<long long> st(B, 7);
  float* x; long long* start; long long* out;
  cudaMalloc(&x, h.size() * 4); cudaMalloc(&start, B * 8); cudaMalloc(&out, (size_t)B * G * 8);
  cudaMemcpy(x, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(start, st.data(), B * 8, cudaMemcpyHostToDevice);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int rep = 0; rep < 6; ++rep) {
    cudaEventRecord(e0);
    int rc = p3tok_fps(x, B, N, 3, (const int64_t*)start, G, (int64_t*)out, nullptr);
    cudaEventRecord(e1);
    cudaError_t e = cudaDeviceSynchronize();
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
    printf("rc=%d cuda=%s  %d clouds x %d pts -> %d: %.3f ms = %.3f us per iteration\n", rc, cudaGetErrorString(e), B, N, G, ms, 1e3 * ms / G);
  }
  // time against the number of clouds: a step up = a second wave (the device cannot seat that many clusters at once)
  for (int b = 1; b <= B; b += (b < 8 ? 3 : 1)) {
    cudaEventRecord(e0);
    p3tok_fps(x, b, N, 3, (const int64_t*)start, G, (int64_t*)out, nullptr);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
    printf("  B = %2d clouds: %.3f ms\n", b, ms);
  }
  std::vector<long long> tr(16 * 32);
  cudaMemcpyFromSymbol(tr.data(), p3tok::fps_trace_buf, tr.size() * 8);
  const int cl = 16;
  printf("iteration 100 of cloud 0, cycles relative to CTA's own warp-0 loop top (per-SM clocks):\n");
  for (int r = 0; r < cl; ++r) {
    const long long t0 = tr[r * 32];
    if (!t0) continue;
    printf("cta %2d warp0:", r);
    for (int i = 0; i < 7; ++i) printf(" t%d=%6lld", i, tr[r * 32 + i] ? tr[r * 32 + i] - t0 : -1);
    printf("   warp5:");
    for (int i = 0; i < 7; ++i) if (tr[r * 32 + 16 + i]) printf(" t%d=%6lld", i, tr[r * 32 + 16 + i] - t0);
    if (tr[r * 32 + 9]) printf("   avg cycles per iteration over g = 200..1200: %.0f", (tr[r * 32 + 9] - tr[r * 32 + 8]) / 1000.0);
    printf("\n");
  }
  std::vector<long long> o((size_t)B * G);
  cudaMemcpy(o.data(), out, o.size() * 8, cudaMemcpyDeviceToHost);
  long long cs = 0; for (auto v : o) cs = cs * 31 + v;
  printf("checksum of picks %lld\n", cs);
  return 0;
}

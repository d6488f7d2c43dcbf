// fps_nd.cu - farthest point sampling on D-dimensional points, 1 <= D <= 16 (sm_100a).
//
// Replaces farthest_point_sampling (reference src/models/pix4point.py:8-53) for inputs whose last dimension is not xyz:
// its distance is torch.sum((points - centroid) ** 2, dim=2) over ALL D coordinates (line 44).  The reference's only call
// site passes xyz (pix4point.py:175) and that case runs on fps_kernel (fps.cu: registers, clusters); this kernel is the
// general-D path of the same function, built to be bit-exact rather than fast.
//
// Arithmetic contract.  Each difference and each square is an individually rounded fp32 operation; the D squares are added
// in the order torch's CPU sum kernel adds a contiguous last dimension (cascade_sum in aten/src/ATen/native/cpu/SumKernel.cpp,
// vector width 8; restated and checked bit for bit against torch.sum for every D in oracle/p3tok_oracle.c:
// torch_cpu_row_sum):
//   D <  8: four interleaved partial sums p[j] = q[j] (j < 4, when D >= 4); the elements from 4*(D/4) on are added to p[0];
//           then p[0] += p[1]; p[0] += p[2]; p[0] += p[3].          (D <= 4: the plain left-to-right sum)
//   D >= 8: lane[k] = q[k] + q[8+k] + ... over the D/8 whole vectors; f = sum of the tail elements (left to right, from 0);
//           then f += lane[0], ..., lane[7].                         (D = 8: the plain left-to-right sum)
// For D <= 3 this is ((dx*dx)+(dy*dy))+(dz*dz): the picks equal fps_kernel's (tests/test_gpu_xtra_index.py cross-checks).
// Running distance starts at 1e10 (pix4point.py:27), update where dist < distance (47-48), argmax keeps the LOWEST index of
// the maximum (torch.max, 51).
//
// Design: one CTA per cloud; the running minima live in a caller-provided (B,N) fp32 scratch (coalesced, L2-resident), the
// points are read in place from global memory every iteration (N*D*4 bytes per iteration and cloud from L2), the argmax is
// thread-local -> redux.sync per warp -> one shared-memory slot per warp -> every warp reduces the slots redundantly, so an
// iteration has ONE __syncthreads (slots double-buffered by iteration parity).  Bound: the chain of L2 round trips of an
// iteration (0.7-1.3 us; the xyz kernel with everything in registers: 0.3-0.6); no BASELINE config reaches this path.
#include "common.cuh"

namespace p3tok {

template <int D>
__device__ __forceinline__ float torch_row_sum(const float (&q)[D]) {
  if constexpr (D < 8) {
    float p[4] = {0.f, 0.f, 0.f, 0.f};
    constexpr int done = D >= 4 ? 4 : 0;
    if constexpr (D >= 4) {
#pragma unroll
      for (int j = 0; j < 4; ++j) p[j] = __fadd_rn(p[j], q[j]);
    }
#pragma unroll
    for (int i = done; i < D; ++i) p[0] = __fadd_rn(p[0], q[i]);
#pragma unroll
    for (int j = 1; j < 4; ++j) p[0] = __fadd_rn(p[0], p[j]);
    return p[0];
  } else {
    constexpr int nvec = D / 8;
    float lane[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) lane[k] = __fadd_rn(0.f, q[k]);
#pragma unroll
    for (int v = 1; v < nvec; ++v) {
#pragma unroll
      for (int k = 0; k < 8; ++k) lane[k] = __fadd_rn(lane[k], q[8 * v + k]);
    }
    float f = 0.f;
#pragma unroll
    for (int i = nvec * 8; i < D; ++i) f = __fadd_rn(f, q[i]);
#pragma unroll
    for (int k = 0; k < 8; ++k) f = __fadd_rn(f, lane[k]);
    return f;
  }
}

// one point's D coordinates -> registers; V4: rows are 16-byte aligned (D % 4 == 0, stride % 4 == 0): 128-bit loads
template <int D, bool V4>
__device__ __forceinline__ void load_point(const float* __restrict__ pr, float (&v)[D]) {
  if constexpr (V4) {
#pragma unroll
    for (int a = 0; a < D; a += 4) {
      const float4 t = __ldg(reinterpret_cast<const float4*>(pr + a));
      v[a] = t.x; v[a + 1] = t.y; v[a + 2] = t.z; v[a + 3] = t.w;
    }
  } else {
#pragma unroll
    for (int a = 0; a < D; ++a) v[a] = __ldg(pr + a);
  }
}

// Two bodies for the pass over the points, chosen by the launch from the measured cross-over (B200, profiles/
// r02_xtra_time*.txt).  STAGED = false: one point at a time - 0.74 us per iteration at N = 1024 (D = 4), 1.11 at N = 2048
// (D = 6), where the whole iteration is one chain of L2 round trips (centroid, minima) and nothing is gained by staging
// (the staged body measured 0.96 / 1.32 there).  STAGED = true, beyond 4 points per thread: U points (4 for D <= 4, 2 for
// D <= 8, else 1) are loaded - 128-bit loads when the rows allow - before any of them is evaluated, so their round trips
// overlap: N = 8192, D = 8: 9.44 -> 4.04 us per iteration.
template <int D, bool V4, bool STAGED>
__global__ void __launch_bounds__(1024)
fps_nd_kernel(const float* __restrict__ x, int N, int64_t pt_stride, const int64_t* __restrict__ start_idx, int G,
              int64_t* __restrict__ out_idx, float* __restrict__ min_dist) {
  constexpr int U = D <= 4 ? 4 : (D <= 8 ? 2 : 1);
  __shared__ uint2 wslot[2][32];
  const int cloud = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nthreads = blockDim.x, nwarps = nthreads >> 5;
  const float* P = x + (int64_t)cloud * N * pt_stride;
  float* md = min_dist + (int64_t)cloud * N;
  int64_t* out = out_idx + (int64_t)cloud * G;

  for (int i = tid; i < N; i += nthreads) md[i] = 1e10f;   // each thread only ever touches its own entries: no barrier needed
  long long s0 = start_idx[cloud];
  int far = (int)(s0 < 0 ? 0 : (s0 >= N ? N - 1 : s0));    // the C ABI cannot validate device data: clamp into the cloud

  for (int g = 0; g < G; ++g) {
    if (tid == 0) out[g] = far;
    float c[D];
    load_point<D, V4>(P + (int64_t)far * pt_stride, c);

    float bm = -1.f;
    uint32_t bi = 0xffffffffu;
    if constexpr (!STAGED) {
      for (int i = tid; i < N; i += nthreads) {
        const float* pr = P + (int64_t)i * pt_stride;
        float q[D];
#pragma unroll
        for (int a = 0; a < D; ++a) {
          const float df = __fsub_rn(__ldg(pr + a), c[a]);
          q[a] = __fmul_rn(df, df);
        }
        const float d = torch_row_sum<D>(q);
        float m = md[i];
        if (d < m) { m = d; md[i] = m; }
        if (m > bm) { bm = m; bi = (uint32_t)i; }             // ascending i, strict >: the thread keeps its lowest index
      }
    } else {
      for (int i0 = tid; i0 < N; i0 += U * nthreads) {
        float v[U][D], m[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int i = i0 + u * nthreads;
          if (i < N) {
            load_point<D, V4>(P + (int64_t)i * pt_stride, v[u]);
            m[u] = md[i];
          }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int i = i0 + u * nthreads;
          if (i < N) {
            float q[D];
#pragma unroll
            for (int a = 0; a < D; ++a) {
              const float df = __fsub_rn(v[u][a], c[a]);
              q[a] = __fmul_rn(df, df);
            }
            const float d = torch_row_sum<D>(q);
            float mm = m[u];
            if (d < mm) { mm = d; md[i] = mm; }
            if (mm > bm) { bm = mm; bi = (uint32_t)i; }       // ascending i, strict >: the thread keeps its lowest index
          }
        }
      }
    }
    // a thread without a pick (no point, or only NaN minima) never wins: key 0 is below f2ord of any value >= -1
    const uint32_t key = bi != 0xffffffffu ? f2ord(bm) : 0u;
    const uint32_t wmax = __reduce_max_sync(0xffffffffu, key);
    const uint32_t widx = __reduce_min_sync(0xffffffffu, key == wmax ? bi : 0xffffffffu);
    const int buf = g & 1;
    if (lane == 0) wslot[buf][warp] = make_uint2(wmax, widx);
    __syncthreads();
    const uint2 s = lane < nwarps ? wslot[buf][lane] : make_uint2(0u, 0xffffffffu);
    const uint32_t bmax = __reduce_max_sync(0xffffffffu, s.x);
    const uint32_t bidx = __reduce_min_sync(0xffffffffu, s.x == bmax ? s.y : 0xffffffffu);
    far = bidx < (uint32_t)N ? (int)bidx : 0;               // all-NaN cloud: stay in range
  }
}

template <int D>
static int fps_nd_launch(const float* x, int B, int N, int64_t pt_stride, const int64_t* start, int G, int64_t* out,
                         float* min_dist, cudaStream_t s) {
  int threads = ((N + 3) / 4 + 31) / 32 * 32;             // 4 points per thread up to 1024 threads (1 or 2 measured no better)
  if (threads > 1024) threads = 1024;
  if (threads < 32) threads = 32;
  if (N <= 4 * 1024) {
    fps_nd_kernel<D, false, false><<<B, threads, 0, s>>>(x, N, pt_stride, start, G, out, min_dist);
  } else if (D % 4 == 0 && pt_stride % 4 == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0) {
    fps_nd_kernel<D, D % 4 == 0, true><<<B, threads, 0, s>>>(x, N, pt_stride, start, G, out, min_dist);
  } else {
    fps_nd_kernel<D, false, true><<<B, threads, 0, s>>>(x, N, pt_stride, start, G, out, min_dist);
  }
  P3_LAUNCH_CHECK("fps_nd_kernel");
  return P3TOK_OK;
}

}  // namespace p3tok

using namespace p3tok;

extern "C" int p3tok_fps_nd(const float* x, int64_t B, int64_t N, int64_t D, int64_t pt_stride, const int64_t* start_idx,
                            int64_t G, int64_t* out_idx, float* min_dist_ws, void* stream) {
  P3_REQUIRE(B >= 0 && N > 0 && G >= 0 && D >= 1 && pt_stride >= D, P3TOK_ERR_INVALID,
             "fps_nd: bad shape B=%lld N=%lld D=%lld G=%lld stride=%lld", (long long)B, (long long)N, (long long)D,
             (long long)G, (long long)pt_stride);
  P3_REQUIRE(D <= 16, P3TOK_ERR_UNSUPPORTED, "fps_nd: D=%lld > 16", (long long)D);
  if (B == 0 || G == 0) return P3TOK_OK;
  P3_REQUIRE(x && start_idx && out_idx && min_dist_ws, P3TOK_ERR_INVALID, "fps_nd: null pointer");
  P3_REQUIRE(N < (1ll << 31) - 8192 && B < (1ll << 31) && G < (1ll << 31), P3TOK_ERR_UNSUPPORTED, "fps_nd: shape too large");
  cudaStream_t s = as_stream(stream);
  const int Bi = (int)B, Ni = (int)N, Gi = (int)G;
  switch ((int)D) {
#define P3_FPS_ND_CASE(d) \
  case d: return fps_nd_launch<d>(x, Bi, Ni, pt_stride, start_idx, Gi, out_idx, min_dist_ws, s);
    P3_FPS_ND_CASE(1) P3_FPS_ND_CASE(2) P3_FPS_ND_CASE(3) P3_FPS_ND_CASE(4) P3_FPS_ND_CASE(5) P3_FPS_ND_CASE(6)
    P3_FPS_ND_CASE(7) P3_FPS_ND_CASE(8) P3_FPS_ND_CASE(9) P3_FPS_ND_CASE(10) P3_FPS_ND_CASE(11) P3_FPS_ND_CASE(12)
    P3_FPS_ND_CASE(13) P3_FPS_ND_CASE(14) P3_FPS_ND_CASE(15) P3_FPS_ND_CASE(16)
#undef P3_FPS_ND_CASE
  }
  return P3TOK_ERR_UNSUPPORTED;
}

// group.cu - gathers, APF grouping and the Morton (Z-order) permutation of patch centres.
//
// Replaces index_points (reference src/data/sampler.py:77-94), the gather / centre-subtract /
// concat / Morton re-ordering of Group.forward (src/models/apf.py:70-110), the gather half of
// group_knn (src/models/pix4point.py:92-102) and MortonEncoder.points_to_morton
// (src/models/apf_utils.py:66-104).  All HBM-bound byte movers: one thread per output element,
// consecutive threads write consecutive addresses.
#include "common.cuh"

namespace p3tok {

__global__ void gather_points_kernel(const float* __restrict__ x, int64_t N, int C,
                                     const int64_t* __restrict__ idx, int64_t S, int64_t total,
                                     float* __restrict__ out) {
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total;
       e += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(e % C);
    const int64_t bs = e / C;
    const int64_t b = bs / S;
    out[e] = x[(b * N + idx[bs]) * C + c];
  }
}

// neigh (B,G,k,2C): thread per element; output group j reads input group perm[b,j]
__global__ void apf_group_kernel(const float* __restrict__ x, int64_t N, int C,
                                 const int64_t* __restrict__ fps_idx, const int64_t* __restrict__ knn_idx,
                                 const int64_t* __restrict__ perm, int64_t G, int k, int64_t total,
                                 float* __restrict__ neigh, float* __restrict__ center) {
  const int C2 = 2 * C;
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total;
       e += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(e % C2);
    const int64_t r = e / C2;          // (b*G + j)*k + n
    const int n = (int)(r % k);
    const int64_t bj = r / k;
    const int64_t b = bj / G;
    const int64_t g = perm ? perm[bj] : (bj - b * G);
    const int64_t ci = fps_idx[b * G + g];
    const float* crow = x + (b * N + ci) * C;
    float v;
    if (c < C) {
      const int64_t ni = knn_idx[(b * G + g) * k + n];
      v = __fsub_rn(x[(b * N + ni) * C + c], crow[c]);   // apf.py:83-84
    } else {
      v = crow[c - C];                                    // apf.py:88-95
    }
    neigh[e] = v;
    if (n == 0 && c < 3) center[bj * 3 + c] = crow[c];
  }
}

__global__ void group_gather_kernel(const float* __restrict__ pnts, const float* __restrict__ feats,
                                    int64_t N, int D, const int32_t* __restrict__ idx, int64_t G, int k,
                                    int64_t rows, float* __restrict__ gp, float* __restrict__ gf) {
  const int W = 3 + D;
  const int64_t total = rows * W;
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total;
       e += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(e % W);
    const int64_t r = e / W;           // (b*G+g)*k + n
    const int64_t b = r / (G * k);
    const int64_t ni = idx[r];
    if (c < 3) gp[r * 3 + c] = pnts[(b * N + ni) * 3 + c];
    else gf[r * D + (c - 3)] = feats[(b * N + ni) * D + (c - 3)];
  }
}

__device__ __forceinline__ int64_t part1by2(int64_t n) {
  n = n & 0x000003ff;
  n = (n ^ (n << 16)) & 0xff0000ff;
  n = (n ^ (n << 8)) & 0x0300f00f;
  n = (n ^ (n << 4)) & 0x030c30c3;
  n = (n ^ (n << 2)) & 0x09249249;
  return n;
}

// one CTA per cloud; keys (code << 32 | g) bitonic-sorted in shared memory (stable by construction)
__global__ void __launch_bounds__(1024)
morton_kernel(const float* __restrict__ centres, int G, int P2, int64_t* __restrict__ perm,
              int64_t* __restrict__ codes_out) {
  extern __shared__ uint64_t keys[];
  __shared__ float red[2][3][32];
  __shared__ float mnmx[2][3];
  const int b = blockIdx.x, t = threadIdx.x, lane = t & 31, warp = t >> 5, nw = blockDim.x >> 5;
  const float* Cn = centres + (size_t)b * G * 3;
  float mn[3] = {3.4e38f, 3.4e38f, 3.4e38f}, mx[3] = {-3.4e38f, -3.4e38f, -3.4e38f};
  for (int g = t; g < G; g += blockDim.x)
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      const float v = Cn[g * 3 + a];
      mn[a] = fminf(mn[a], v);
      mx[a] = fmaxf(mx[a], v);
    }
#pragma unroll
  for (int a = 0; a < 3; ++a) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      mn[a] = fminf(mn[a], __shfl_xor_sync(0xffffffffu, mn[a], o));
      mx[a] = fmaxf(mx[a], __shfl_xor_sync(0xffffffffu, mx[a], o));
    }
    if (lane == 0) { red[0][a][warp] = mn[a]; red[1][a][warp] = mx[a]; }
  }
  __syncthreads();
  if (warp == 0) {
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      float m0 = lane < nw ? red[0][a][lane] : 3.4e38f, m1 = lane < nw ? red[1][a][lane] : -3.4e38f;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        m0 = fminf(m0, __shfl_xor_sync(0xffffffffu, m0, o));
        m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, o));
      }
      if (lane == 0) { mnmx[0][a] = m0; mnmx[1][a] = m1; }
    }
  }
  __syncthreads();
  for (int g = t; g < P2; g += blockDim.x) {
    uint64_t key = 0xffffffffffffffffull;
    if (g < G) {
      int64_t q[3];
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        const float num = __fsub_rn(Cn[g * 3 + a], mnmx[0][a]);
        const float den = __fadd_rn(__fsub_rn(mnmx[1][a], mnmx[0][a]), 1e-8f);   // apf_utils.py:91
        const float sc = __fmul_rn(__fdiv_rn(num, den), 1023.0f);               // apf_utils.py:92
        q[a] = (int64_t)sc;                                                      // .long(): truncation
      }
      const int64_t code = (part1by2(q[2]) << 2) + (part1by2(q[1]) << 1) + part1by2(q[0]);
      if (codes_out) codes_out[(size_t)b * G + g] = code;
      key = ((uint64_t)code << 32) | (uint32_t)g;
    }
    keys[g] = key;
  }
  __syncthreads();
  for (int k = 2; k <= P2; k <<= 1)
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = t; i < P2; i += blockDim.x) {
        const int p = i ^ j;
        if (p > i) {
          const uint64_t a = keys[i], c = keys[p];
          const bool up = (i & k) == 0;
          if ((a > c) == up) { keys[i] = c; keys[p] = a; }
        }
      }
      __syncthreads();
    }
  for (int g = t; g < G; g += blockDim.x) perm[(size_t)b * G + g] = (int64_t)(keys[g] & 0xffffffffu);
}

static inline unsigned blocks_for(int64_t total, int threads) {
  int64_t b = (total + threads - 1) / threads;
  const int64_t cap = 148 * 16;
  return (unsigned)(b < 1 ? 1 : (b > cap ? cap : b));
}

}  // namespace p3tok

using namespace p3tok;

extern "C" int p3tok_gather_points(const float* x, int64_t B, int64_t N, int64_t C, const int64_t* idx,
                                   int64_t S, float* out, void* stream) {
  P3_REQUIRE(B >= 0 && N > 0 && C > 0 && S >= 0, P3TOK_ERR_INVALID, "gather_points: bad shape");
  const int64_t total = B * S * C;
  if (total == 0) return P3TOK_OK;
  P3_REQUIRE(x && idx && out, P3TOK_ERR_INVALID, "gather_points: null pointer");
  gather_points_kernel<<<blocks_for(total, 256), 256, 0, as_stream(stream)>>>(x, N, (int)C, idx, S, total, out);
  P3_LAUNCH_CHECK("gather_points_kernel");
  return P3TOK_OK;
}

extern "C" int p3tok_apf_group(const float* x, int64_t B, int64_t N, int64_t C, const int64_t* fps_idx,
                               const int64_t* knn_idx, const int64_t* perm, int64_t G, int64_t k,
                               float* neigh, float* center, void* stream) {
  P3_REQUIRE(B >= 0 && N > 0 && C >= 3 && G >= 0 && k > 0, P3TOK_ERR_INVALID, "apf_group: bad shape");
  const int64_t total = B * G * k * 2 * C;
  if (total == 0) return P3TOK_OK;
  P3_REQUIRE(x && fps_idx && knn_idx && neigh && center, P3TOK_ERR_INVALID, "apf_group: null pointer");
  apf_group_kernel<<<blocks_for(total, 256), 256, 0, as_stream(stream)>>>(x, N, (int)C, fps_idx, knn_idx, perm, G,
                                                                          (int)k, total, neigh, center);
  P3_LAUNCH_CHECK("apf_group_kernel");
  return P3TOK_OK;
}

extern "C" int p3tok_group_gather(const float* pnts, const float* feats, int64_t B, int64_t N, int64_t D,
                                  const int32_t* idx, int64_t G, int64_t k, float* gp, float* gf, void* stream) {
  P3_REQUIRE(B >= 0 && N > 0 && D >= 0 && G >= 0 && k > 0, P3TOK_ERR_INVALID, "group_gather: bad shape");
  const int64_t rows = B * G * k;
  if (rows == 0) return P3TOK_OK;
  P3_REQUIRE(pnts && idx && gp && (D == 0 || (feats && gf)), P3TOK_ERR_INVALID, "group_gather: null pointer");
  group_gather_kernel<<<blocks_for(rows * (3 + D), 256), 256, 0, as_stream(stream)>>>(pnts, feats, N, (int)D, idx, G,
                                                                                     (int)k, rows, gp, gf);
  P3_LAUNCH_CHECK("group_gather_kernel");
  return P3TOK_OK;
}

extern "C" int p3tok_morton_order(const float* centres, int64_t B, int64_t G, int64_t* perm, int64_t* codes_out,
                                  void* stream) {
  P3_REQUIRE(B >= 0 && G > 0, P3TOK_ERR_INVALID, "morton_order: bad shape");
  P3_REQUIRE(G <= 8192, P3TOK_ERR_UNSUPPORTED, "morton_order: G=%lld > 8192", (long long)G);
  if (B == 0) return P3TOK_OK;
  P3_REQUIRE(centres && perm, P3TOK_ERR_INVALID, "morton_order: null pointer");
  int P2 = 1;
  while (P2 < G) P2 <<= 1;
  int threads = P2 < 64 ? 64 : (P2 > 1024 ? 1024 : P2);
  const size_t smem = (size_t)P2 * 8;
  if (smem > 48 * 1024) P3_CUDA(cudaFuncSetAttribute(morton_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
  morton_kernel<<<(unsigned)B, threads, smem, as_stream(stream)>>>(centres, (int)G, P2, perm, codes_out);
  P3_LAUNCH_CHECK("morton_kernel");
  return P3TOK_OK;
}

// vit.cu - the ViT block stack that consumes the tokens (SURVEY.md 8f "next" #3).
//
// Replaces the block loop, encoder_norm and token max of AdaptPointFormer.forward (reference src/models/apf.py:361-366)
// with APFViTLayer = attention + bottleneck adapter + MLP (src/models/apf_utils.py:106-293), eval mode.
//
// Data flow per layer (M = B*G token rows, fp32 residual stream x updated in place, bf16 activations in the workspace):
//   ln_rows_kernel        a  = bf16(norm1(x))                                           warp per row
//   tc_linear             qkv = a Wqkv^T + b                   (tcgen05, bf16 out)      embed_tc.cu
//   attention_kernel      o  = softmax(q k^T / sqrt(hd)) v     per (cloud, head, 64 query rows)
//   tc_linear             x  = x + (o Wproj^T + b)             (fp32 residual epilogue)
//   ln_rows_kernel        n2 = bf16(norm2(x)),  an = bf16(adapter_norm(x))              one pass, shared statistics
//   tc_linear             dn = relu(an Wdown^T + b)
//   tc_linear             x  = 2 x + scale (dn Wup^T + b)      (adapter output + the layer's own residual, as the
//                                                               reference adds them: apf_utils.py:233 and :292)
//   tc_linear             h  = gelu(n2 Wfc1^T + b)             (exact GELU in the epilogue)
//   tc_linear             x  = x + (h Wfc2^T + b)
// then norm_max_kernel: pooled[b] = max over tokens of encoder_norm(x[b]).
//
// The GEMMs (95 % of the FLOPs) run on the tcgen05 kernel of embed_tc.cu; attention (sequence 128-196, head dim 32/64:
// 5 % of the FLOPs, exp-bound) is a flash-style kernel on warp-level bf16 mma with fp32 online softmax, K and V^T of
// one (cloud, head) staged in shared memory per 64-key block.
#include <math.h>

#include "embed.cuh"

namespace p3tok {

// ------------------------------------------------------------------------------------------------ LayerNorm
// One warp per row, the row in registers (float4 index lane + 32 i, D <= 1024), two passes (mean, then centred
// variance - what torch computes up to rounding), biased variance, eps inside the square root.
constexpr int LN_MAX_V4 = 8;

template <bool DUAL>
__global__ void __launch_bounds__(256)
ln_rows_kernel(const float* __restrict__ x, int64_t M, int D, float eps, const float* __restrict__ w1,
               const float* __restrict__ b1, __nv_bfloat16* __restrict__ o1, const float* __restrict__ w2,
               const float* __restrict__ b2, __nv_bfloat16* __restrict__ o2) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= M) return;
  const int nv = D >> 2;
  const float4* xr = reinterpret_cast<const float4*>(x + row * D);
  float4 v[LN_MAX_V4];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < LN_MAX_V4; ++i) {
    const int j = lane + 32 * i;
    v[i] = j < nv ? xr[j] : make_float4(0.f, 0.f, 0.f, 0.f);
    s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float mean = s / (float)D;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < LN_MAX_V4; ++i) {
    const int j = lane + 32 * i;
    if (j < nv) {
      const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
      q += (a * a + b * b) + (c * c + d * d);
    }
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
  const float rstd = 1.f / sqrtf(q / (float)D + eps);
#pragma unroll
  for (int i = 0; i < LN_MAX_V4; ++i) {
    const int j = lane + 32 * i;
    if (j < nv) {
      const float a = (v[i].x - mean) * rstd, b = (v[i].y - mean) * rstd, c = (v[i].z - mean) * rstd,
                  d = (v[i].w - mean) * rstd;
      {
        const float4 w = reinterpret_cast<const float4*>(w1)[j], bb = reinterpret_cast<const float4*>(b1)[j];
        __nv_bfloat162 lo = __floats2bfloat162_rn(fmaf(a, w.x, bb.x), fmaf(b, w.y, bb.y));
        __nv_bfloat162 hi = __floats2bfloat162_rn(fmaf(c, w.z, bb.z), fmaf(d, w.w, bb.w));
        uint2 pk = make_uint2(*reinterpret_cast<uint32_t*>(&lo), *reinterpret_cast<uint32_t*>(&hi));
        reinterpret_cast<uint2*>(o1 + row * D)[j] = pk;
      }
      if (DUAL) {
        const float4 w = reinterpret_cast<const float4*>(w2)[j], bb = reinterpret_cast<const float4*>(b2)[j];
        __nv_bfloat162 lo = __floats2bfloat162_rn(fmaf(a, w.x, bb.x), fmaf(b, w.y, bb.y));
        __nv_bfloat162 hi = __floats2bfloat162_rn(fmaf(c, w.z, bb.z), fmaf(d, w.w, bb.w));
        uint2 pk = make_uint2(*reinterpret_cast<uint32_t*>(&lo), *reinterpret_cast<uint32_t*>(&hi));
        reinterpret_cast<uint2*>(o2 + row * D)[j] = pk;
      }
    }
  }
}

static int layernorm_bf16(const float* x, int64_t M, int D, float eps, const float* w1, const float* b1, __nv_bfloat16* o1,
                          const float* w2, const float* b2, __nv_bfloat16* o2, cudaStream_t s) {
  P3_REQUIRE(D % 4 == 0 && D > 0 && D <= 128 * LN_MAX_V4, P3TOK_ERR_UNSUPPORTED, "layernorm: D=%d must be a multiple of 4, <= %d", D,
             128 * LN_MAX_V4);
  P3_REQUIRE(x && w1 && b1 && o1, P3TOK_ERR_INVALID, "layernorm: null pointer");
  if (M == 0) return P3TOK_OK;
  const unsigned blocks = (unsigned)((M + 7) / 8);
  if (o2) {
    P3_REQUIRE(w2 && b2, P3TOK_ERR_INVALID, "layernorm: second affine missing");
    ln_rows_kernel<true><<<blocks, 256, 0, s>>>(x, M, D, eps, w1, b1, o1, w2, b2, o2);
  } else {
    ln_rows_kernel<false><<<blocks, 256, 0, s>>>(x, M, D, eps, w1, b1, o1, nullptr, nullptr, nullptr);
  }
  P3_LAUNCH_CHECK("ln_rows_kernel");
  return P3TOK_OK;
}

// encoder_norm + max over the tokens of a cloud (apf.py:364-366).  One CTA per cloud, warp per token row.
__global__ void __launch_bounds__(256)
norm_max_kernel(const float* __restrict__ x, int G, int D, float eps, const float* __restrict__ w, const float* __restrict__ b,
                float* __restrict__ pooled) {
  __shared__ float4 red[8][32 * LN_MAX_V4];   // [warp][D/4], 32 KB
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nv = D >> 2;
  const float* xb = x + (size_t)blockIdx.x * G * D;
  float4 mx[LN_MAX_V4];
#pragma unroll
  for (int i = 0; i < LN_MAX_V4; ++i) mx[i] = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
  for (int r = warp; r < G; r += 8) {
    const float4* xr = reinterpret_cast<const float4*>(xb + (size_t)r * D);
    float4 v[LN_MAX_V4];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < LN_MAX_V4; ++i) {
      const int j = lane + 32 * i;
      v[i] = j < nv ? xr[j] : make_float4(0.f, 0.f, 0.f, 0.f);
      s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s / (float)D;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < LN_MAX_V4; ++i) {
      const int j = lane + 32 * i;
      if (j < nv) {
        const float a = v[i].x - mean, bq = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
        q += (a * a + bq * bq) + (c * c + d * d);
      }
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    const float rstd = 1.f / sqrtf(q / (float)D + eps);
#pragma unroll
    for (int i = 0; i < LN_MAX_V4; ++i) {
      const int j = lane + 32 * i;
      if (j < nv) {
        const float4 ww = reinterpret_cast<const float4*>(w)[j], bb = reinterpret_cast<const float4*>(b)[j];
        mx[i].x = fmaxf(mx[i].x, fmaf((v[i].x - mean) * rstd, ww.x, bb.x));
        mx[i].y = fmaxf(mx[i].y, fmaf((v[i].y - mean) * rstd, ww.y, bb.y));
        mx[i].z = fmaxf(mx[i].z, fmaf((v[i].z - mean) * rstd, ww.z, bb.z));
        mx[i].w = fmaxf(mx[i].w, fmaf((v[i].w - mean) * rstd, ww.w, bb.w));
      }
    }
  }
#pragma unroll
  for (int i = 0; i < LN_MAX_V4; ++i) red[warp][lane + 32 * i] = mx[i];
  __syncthreads();
  for (int j = threadIdx.x; j < nv; j += blockDim.x) {
    float4 m = red[0][j];
#pragma unroll
    for (int wv = 1; wv < 8; ++wv) {
      const float4 o = red[wv][j];
      m.x = fmaxf(m.x, o.x); m.y = fmaxf(m.y, o.y); m.z = fmaxf(m.z, o.z); m.w = fmaxf(m.w, o.w);
    }
    reinterpret_cast<float4*>(pooled + (size_t)blockIdx.x * D)[j] = m;
  }
}

// ------------------------------------------------------------------------------------------------ attention
__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack2(float a, float b) {
  __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&t);
}

constexpr int AT_Q = 64, AT_KV = 64;   // query rows per CTA (4 warps x 16), keys per block

// qkv (B*G, 3D) bf16: column which*D + head*HD + d (AttentionLayer's reshape(B,N,3,heads,hd), apf_utils.py:143).
// out (B*G, D) bf16: column head*HD + d ((attn @ v).transpose(1,2).reshape(B,N,C), apf_utils.py:155).
// grid (ceil(G/64), heads, B), 128 threads.  Lane (g = lane/4, t = lane%4) of a warp owns query rows g and g+8 of
// the warp's 16: the m16n8k16 accumulator layout, so row maxima / sums need only the two quad shuffles.
template <int HD>
__global__ void __launch_bounds__(128)
attention_kernel(const __nv_bfloat16* __restrict__ qkv, __nv_bfloat16* __restrict__ out, int G, int D, float scale_log2e) {
  constexpr int KS = HD / 16;          // k steps of q k^T
  constexpr int DT = HD / 8;           // n tiles of the output
  constexpr int KP = HD + 8;           // padded row of sK (conflict-free fragment loads)
  constexpr int VP = AT_KV + 8;        // padded row of sVt
  __shared__ __align__(16) __nv_bfloat16 sK[AT_KV * KP];
  __shared__ __align__(16) __nv_bfloat16 sVt[HD * VP];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int head = blockIdx.y, b = blockIdx.z;
  const int q0 = blockIdx.x * AT_Q + warp * 16;
  const size_t ld = (size_t)3 * D;
  const __nv_bfloat16* base = qkv + (size_t)b * G * ld + (size_t)head * HD;

  // Q fragments straight from global memory (rows past G are clamped; their results are never stored)
  uint32_t qa[KS][4];
  {
    const int r0 = min(q0 + g, G - 1), r1 = min(q0 + g + 8, G - 1);
    const __nv_bfloat16* p0 = base + (size_t)r0 * ld;
    const __nv_bfloat16* p1 = base + (size_t)r1 * ld;
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
      qa[ks][0] = *reinterpret_cast<const uint32_t*>(p0 + ks * 16 + 2 * t);
      qa[ks][1] = *reinterpret_cast<const uint32_t*>(p1 + ks * 16 + 2 * t);
      qa[ks][2] = *reinterpret_cast<const uint32_t*>(p0 + ks * 16 + 8 + 2 * t);
      qa[ks][3] = *reinterpret_cast<const uint32_t*>(p1 + ks * 16 + 8 + 2 * t);
    }
  }
  float o[DT][4];
#pragma unroll
  for (int i = 0; i < DT; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
  float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;

  for (int kv0 = 0; kv0 < G; kv0 += AT_KV) {
    __syncthreads();   // previous block's fragments have been read
    for (int c = tid; c < AT_KV * (HD / 8); c += 128) {
      const int r = c / (HD / 8), ch = c % (HD / 8);
      uint4 kk = make_uint4(0, 0, 0, 0), vv = make_uint4(0, 0, 0, 0);
      if (kv0 + r < G) {
        const __nv_bfloat16* rowp = base + (size_t)(kv0 + r) * ld + ch * 8;
        kk = *reinterpret_cast<const uint4*>(rowp + D);
        vv = *reinterpret_cast<const uint4*>(rowp + 2 * D);
      }
      *reinterpret_cast<uint4*>(&sK[r * KP + ch * 8]) = kk;
      const __nv_bfloat16* ve = reinterpret_cast<const __nv_bfloat16*>(&vv);
#pragma unroll
      for (int i = 0; i < 8; ++i) sVt[(ch * 8 + i) * VP + r] = ve[i];
    }
    __syncthreads();

    float sc[AT_KV / 8][4];
#pragma unroll
    for (int nt = 0; nt < AT_KV / 8; ++nt) {
      sc[nt][0] = sc[nt][1] = sc[nt][2] = sc[nt][3] = 0.f;
#pragma unroll
      for (int ks = 0; ks < KS; ++ks) {
        const uint32_t b0 = *reinterpret_cast<const uint32_t*>(&sK[(nt * 8 + g) * KP + ks * 16 + 2 * t]);
        const uint32_t b1 = *reinterpret_cast<const uint32_t*>(&sK[(nt * 8 + g) * KP + ks * 16 + 8 + 2 * t]);
        mma_bf16_16816(sc[nt], qa[ks], b0, b1);
      }
    }
    // scale (log2 domain), mask keys past G, block row maxima
    float bm0 = -INFINITY, bm1 = -INFINITY;
#pragma unroll
    for (int nt = 0; nt < AT_KV / 8; ++nt) {
      const int col = kv0 + nt * 8 + 2 * t;
      const bool ok0 = col < G, ok1 = col + 1 < G;
      sc[nt][0] = ok0 ? sc[nt][0] * scale_log2e : -INFINITY;
      sc[nt][1] = ok1 ? sc[nt][1] * scale_log2e : -INFINITY;
      sc[nt][2] = ok0 ? sc[nt][2] * scale_log2e : -INFINITY;
      sc[nt][3] = ok1 ? sc[nt][3] * scale_log2e : -INFINITY;
      bm0 = fmaxf(bm0, fmaxf(sc[nt][0], sc[nt][1]));
      bm1 = fmaxf(bm1, fmaxf(sc[nt][2], sc[nt][3]));
    }
    bm0 = fmaxf(bm0, __shfl_xor_sync(0xffffffffu, bm0, 1));
    bm0 = fmaxf(bm0, __shfl_xor_sync(0xffffffffu, bm0, 2));
    bm1 = fmaxf(bm1, __shfl_xor_sync(0xffffffffu, bm1, 1));
    bm1 = fmaxf(bm1, __shfl_xor_sync(0xffffffffu, bm1, 2));
    const float n0 = fmaxf(m0, bm0), n1 = fmaxf(m1, bm1);   // finite: every block holds at least one valid key
    const float c0 = exp2f(m0 - n0), c1 = exp2f(m1 - n1);   // 0 on the first block (m = -inf)
    m0 = n0; m1 = n1;
    l0 *= c0; l1 *= c1;
#pragma unroll
    for (int i = 0; i < DT; ++i) { o[i][0] *= c0; o[i][1] *= c0; o[i][2] *= c1; o[i][3] *= c1; }
#pragma unroll
    for (int nt = 0; nt < AT_KV / 8; ++nt) {
      sc[nt][0] = exp2f(sc[nt][0] - n0); sc[nt][1] = exp2f(sc[nt][1] - n0);
      sc[nt][2] = exp2f(sc[nt][2] - n1); sc[nt][3] = exp2f(sc[nt][3] - n1);
      l0 += sc[nt][0] + sc[nt][1];
      l1 += sc[nt][2] + sc[nt][3];
    }
    // O += P V: the score accumulators of two adjacent key tiles are exactly one A fragment
#pragma unroll
    for (int kk = 0; kk < AT_KV / 16; ++kk) {
      uint32_t pa[4];
      pa[0] = pack2(sc[2 * kk][0], sc[2 * kk][1]);
      pa[1] = pack2(sc[2 * kk][2], sc[2 * kk][3]);
      pa[2] = pack2(sc[2 * kk + 1][0], sc[2 * kk + 1][1]);
      pa[3] = pack2(sc[2 * kk + 1][2], sc[2 * kk + 1][3]);
#pragma unroll
      for (int dt = 0; dt < DT; ++dt) {
        const uint32_t b0 = *reinterpret_cast<const uint32_t*>(&sVt[(dt * 8 + g) * VP + kk * 16 + 2 * t]);
        const uint32_t b1 = *reinterpret_cast<const uint32_t*>(&sVt[(dt * 8 + g) * VP + kk * 16 + 8 + 2 * t]);
        mma_bf16_16816(o[dt], pa, b0, b1);
      }
    }
  }
  l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
  l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
  const float i0 = 1.f / l0, i1 = 1.f / l1;
  const int r0 = q0 + g, r1 = q0 + g + 8;
  __nv_bfloat16* ob = out + (size_t)b * G * D + (size_t)head * HD;
#pragma unroll
  for (int dt = 0; dt < DT; ++dt) {
    if (r0 < G) *reinterpret_cast<uint32_t*>(ob + (size_t)r0 * D + dt * 8 + 2 * t) = pack2(o[dt][0] * i0, o[dt][1] * i0);
    if (r1 < G) *reinterpret_cast<uint32_t*>(ob + (size_t)r1 * D + dt * 8 + 2 * t) = pack2(o[dt][2] * i1, o[dt][3] * i1);
  }
}

static int attention_bf16(const __nv_bfloat16* qkv, int64_t B, int64_t G, int D, int heads, __nv_bfloat16* out, cudaStream_t s) {
  P3_REQUIRE(heads > 0 && D % heads == 0, P3TOK_ERR_INVALID, "attention: D=%d not divisible by heads=%d", D, heads);
  const int hd = D / heads;
  P3_REQUIRE(hd == 32 || hd == 64, P3TOK_ERR_UNSUPPORTED, "attention: head dim %d (supported: 32, 64)", hd);
  P3_REQUIRE(B <= 65535 && heads <= 65535, P3TOK_ERR_UNSUPPORTED, "attention: B=%lld / heads=%d exceed the grid", (long long)B, heads);
  P3_REQUIRE(qkv && out, P3TOK_ERR_INVALID, "attention: null pointer");
  P3_REQUIRE((reinterpret_cast<uintptr_t>(qkv) & 15) == 0, P3TOK_ERR_UNSUPPORTED, "attention: qkv must be 16-byte aligned");
  if (B * G == 0) return P3TOK_OK;
  const float scale_log2e = (float)((1.0 / sqrt((double)hd)) * 1.4426950408889634);
  dim3 grid((unsigned)((G + AT_Q - 1) / AT_Q), (unsigned)heads, (unsigned)B);
  if (hd == 32) attention_kernel<32><<<grid, 128, 0, s>>>(qkv, out, (int)G, D, scale_log2e);
  else attention_kernel<64><<<grid, 128, 0, s>>>(qkv, out, (int)G, D, scale_log2e);
  P3_LAUNCH_CHECK("attention_kernel");
  return P3TOK_OK;
}

// ------------------------------------------------------------------------------------------------ orchestration
struct VitWs {
  int64_t a, a2, qkv, h, dn, total;
};
static VitWs vit_layout(int64_t M, int64_t D, int64_t H, int64_t R) {
  VitWs w;
  int64_t off = 0;
  auto take = [&](int64_t bytes) { int64_t o = off; off += align_up(bytes, 1024); return o; };
  w.a = take(M * D * 2);
  w.a2 = take(M * D * 2);
  w.qkv = take(M * 3 * D * 2);
  w.h = take(M * H * 2);
  w.dn = take(M * R * 2);
  w.total = off + 1024;
  return w;
}
int64_t apf_vit_workspace(int64_t B, int64_t G, int64_t D, int64_t H, int64_t R) { return vit_layout(B * G, D, H, R).total; }

// bf16-output GEMM whose N may exceed what one tc_linear launch stages (2048 columns): column slices of the weight
// matrix write column slices of the output (ViT-B: 3D = 2304, H = 3072)
static int linear_wide(const __nv_bfloat16* A, int64_t M, int K, const __nv_bfloat16* W, int N, const float* bias, int relu, int gelu,
                       __nv_bfloat16* out, cudaStream_t s) {
  const int parts = (N + 2047) / 2048;
  int per = ((N + parts - 1) / parts + 63) / 64 * 64;
  for (int n0 = 0; n0 < N; n0 += per) {
    const int n = N - n0 < per ? N - n0 : per;
    TcExtra ex;
    ex.gelu = gelu;
    ex.ldc = N;
    int rc = tc_linear_ex(A, M, K, W + (size_t)n0 * K, n, bias + n0, relu, ex, out + n0, nullptr, s);
    if (rc) return rc;
  }
  return P3TOK_OK;
}

}  // namespace p3tok

using namespace p3tok;

extern "C" int64_t p3tok_apf_vit_workspace_bytes(int64_t B, int64_t G, int64_t D, int64_t H, int64_t R) {
  if (B < 0 || G < 0 || D <= 0 || H <= 0 || R <= 0) return -1;
  return apf_vit_workspace(B, G, D, H, R);
}

extern "C" int p3tok_layernorm_bf16(const float* x, int64_t M, int64_t D, float eps, const float* w1, const float* b1, void* out1,
                                    const float* w2, const float* b2, void* out2, void* stream) {
  P3_REQUIRE(M >= 0 && D > 0, P3TOK_ERR_INVALID, "layernorm: bad shape");
  return layernorm_bf16(x, M, (int)D, eps, w1, b1, (__nv_bfloat16*)out1, w2, b2, (__nv_bfloat16*)out2, as_stream(stream));
}

extern "C" int p3tok_attention_bf16(const void* qkv, int64_t B, int64_t G, int64_t D, int64_t heads, void* out, void* stream) {
  P3_REQUIRE(B >= 0 && G >= 0 && D > 0 && D % 8 == 0, P3TOK_ERR_INVALID, "attention: bad shape");
  return attention_bf16((const __nv_bfloat16*)qkv, B, G, (int)D, (int)heads, (__nv_bfloat16*)out, as_stream(stream));
}

extern "C" int p3tok_linear_bf16_ex(const void* A, int64_t M, int64_t K, const void* W, int64_t N, const float* bias, int act,
                                    const float* residual, float res_mul, float out_scale, void* out_bf16, float* out_f32,
                                    void* stream) {
  P3_REQUIRE(M >= 0 && K > 0 && N > 0 && K < (1 << 24) && N < (1 << 24), P3TOK_ERR_INVALID, "linear_bf16_ex: bad shape");
  P3_REQUIRE(act >= 0 && act <= 2, P3TOK_ERR_INVALID, "linear_bf16_ex: act %d", act);
  if (M == 0) return P3TOK_OK;
  P3_REQUIRE(A && W && (out_bf16 || out_f32), P3TOK_ERR_INVALID, "linear_bf16_ex: null pointer");
  P3_REQUIRE(!residual || out_f32, P3TOK_ERR_INVALID, "linear_bf16_ex: a residual needs the f32 output");
  TcExtra ex;
  ex.gelu = act == 2;
  ex.residual = residual;
  ex.res_mul = res_mul;
  ex.out_scale = residual ? out_scale : 1.f;
  return tc_linear_ex((const __nv_bfloat16*)A, M, (int)K, (const __nv_bfloat16*)W, (int)N, bias, act == 1, ex,
                      (__nv_bfloat16*)out_bf16, out_f32, as_stream(stream));
}

extern "C" int p3tok_apf_vit_forward(float* x, int64_t B, int64_t G, int64_t D, int64_t heads, int64_t H, int64_t R,
                                     const p3tok_vit_layer* layers, int64_t n_layers, const float* final_norm_w,
                                     const float* final_norm_b, float* pooled_out, void* workspace, int64_t workspace_bytes,
                                     void* stream) {
  P3_REQUIRE(B >= 0 && G >= 0 && D > 0 && H > 0 && R > 0 && heads > 0 && n_layers >= 0, P3TOK_ERR_INVALID, "apf_vit: bad shape");
  P3_REQUIRE(D % 8 == 0 && H % 8 == 0 && R % 8 == 0, P3TOK_ERR_UNSUPPORTED, "apf_vit: D, H, R must be multiples of 8");
  P3_REQUIRE(D <= 128 * LN_MAX_V4, P3TOK_ERR_UNSUPPORTED, "apf_vit: D=%lld > %d", (long long)D, 128 * LN_MAX_V4);
  P3_REQUIRE(D % heads == 0 && (D / heads == 32 || D / heads == 64), P3TOK_ERR_UNSUPPORTED,
             "apf_vit: head dim %lld (supported: 32, 64)", (long long)(D / heads));
  P3_REQUIRE(B * G < (1ll << 31) - 256, P3TOK_ERR_UNSUPPORTED, "apf_vit: too many token rows");
  const int64_t M = B * G;
  if (M == 0) return P3TOK_OK;
  P3_REQUIRE(x && workspace && (n_layers == 0 || layers), P3TOK_ERR_INVALID, "apf_vit: null pointer");
  P3_REQUIRE(!pooled_out || (final_norm_w && final_norm_b), P3TOK_ERR_INVALID, "apf_vit: pooled output needs encoder_norm");
  const VitWs L = vit_layout(M, D, H, R);
  P3_REQUIRE(workspace_bytes >= L.total, P3TOK_ERR_WORKSPACE, "apf_vit: workspace %lld < %lld bytes", (long long)workspace_bytes,
             (long long)L.total);
  cudaStream_t s = as_stream(stream);
  uint8_t* ws = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(workspace) + 1023) & ~(uintptr_t)1023);
  __nv_bfloat16* a = reinterpret_cast<__nv_bfloat16*>(ws + L.a);
  __nv_bfloat16* a2 = reinterpret_cast<__nv_bfloat16*>(ws + L.a2);
  __nv_bfloat16* qkv = reinterpret_cast<__nv_bfloat16*>(ws + L.qkv);
  __nv_bfloat16* h = reinterpret_cast<__nv_bfloat16*>(ws + L.h);
  __nv_bfloat16* dn = reinterpret_cast<__nv_bfloat16*>(ws + L.dn);
  const float eps = 1e-5f;
  int rc;
  for (int64_t li = 0; li < n_layers; ++li) {
    const p3tok_vit_layer& w = layers[li];
    P3_REQUIRE(w.norm1_w && w.norm1_b && w.norm2_w && w.norm2_b && w.adnorm_w && w.adnorm_b && w.qkv_w && w.qkv_b && w.proj_w &&
                   w.proj_b && w.fc1_w && w.fc1_b && w.fc2_w && w.fc2_b && w.down_w && w.down_b && w.up_w && w.up_b,
               P3TOK_ERR_INVALID, "apf_vit: layer %lld has a null parameter", (long long)li);
    // attention branch: x += proj(attention(norm1(x)))                                    (apf_utils.py:279-283)
    if ((rc = layernorm_bf16(x, M, (int)D, eps, w.norm1_w, w.norm1_b, a, nullptr, nullptr, nullptr, s))) return rc;
    if ((rc = linear_wide(a, M, (int)D, (const __nv_bfloat16*)w.qkv_w, (int)(3 * D), w.qkv_b, 0, 0, qkv, s))) return rc;
    if ((rc = attention_bf16(qkv, B, G, (int)D, (int)heads, a, s))) return rc;
    {
      TcExtra ex;
      ex.residual = x; ex.res_mul = 1.f; ex.out_scale = 1.f;
      if ((rc = tc_linear_ex(a, M, (int)D, (const __nv_bfloat16*)w.proj_w, (int)D, w.proj_b, 0, ex, nullptr, x, s))) return rc;
    }
    // adapter + MLP on the same x: out = mlp(norm2(x)) + [scale * up(relu(down(adapter_norm(x)))) + x] + x   (:284-292)
    if ((rc = layernorm_bf16(x, M, (int)D, eps, w.norm2_w, w.norm2_b, a, w.adnorm_w, w.adnorm_b, a2, s))) return rc;
    {
      TcExtra ex;
      if ((rc = tc_linear_ex(a2, M, (int)D, (const __nv_bfloat16*)w.down_w, (int)R, w.down_b, 1, ex, dn, nullptr, s))) return rc;
    }
    {
      TcExtra ex;
      ex.residual = x; ex.res_mul = 2.f; ex.out_scale = w.adapter_scale;
      if ((rc = tc_linear_ex(dn, M, (int)R, (const __nv_bfloat16*)w.up_w, (int)D, w.up_b, 0, ex, nullptr, x, s))) return rc;
    }
    if ((rc = linear_wide(a, M, (int)D, (const __nv_bfloat16*)w.fc1_w, (int)H, w.fc1_b, 0, 1, h, s))) return rc;
    {
      TcExtra ex;
      ex.residual = x; ex.res_mul = 1.f; ex.out_scale = 1.f;
      if ((rc = tc_linear_ex(h, M, (int)H, (const __nv_bfloat16*)w.fc2_w, (int)D, w.fc2_b, 0, ex, nullptr, x, s))) return rc;
    }
  }
  if (pooled_out) {
    norm_max_kernel<<<(unsigned)B, 256, 0, s>>>(x, (int)G, (int)D, eps, final_norm_w, final_norm_b, pooled_out);
    P3_LAUNCH_CHECK("norm_max_kernel");
  }
  return P3TOK_OK;
}

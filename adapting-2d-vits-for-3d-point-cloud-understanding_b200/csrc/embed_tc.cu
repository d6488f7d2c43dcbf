// embed_tc.cu - bf16 patch embedding on the 5th-generation tensor cores (tcgen05 + TMEM + TMA).
//
// Replaces Encoder.get_features (reference src/models/apf.py:145-169) and the conv1/conv2/pool
// half of a P3Embed stage (src/models/pix4point.py:179-188), eval-mode BatchNorm folded on the
// host, in the rtol-1e-2 precision mode: bf16 operands, fp32 accumulation in tensor memory.
//
// Kernels
//   rows_first_layer_kernel  gather + centre-subtract + first layer on CUDA cores when the input
//                            is narrow (cin <= 16: APF 2C = 6/8, P3Embed stage 0 = 6), fp32 math
//                            from fp32 coordinates, bf16 out.  Wide inputs (P3Embed stage 1,
//                            cin = 131) are instead gathered to a zero-padded bf16 row matrix
//                            (rows_gather_bf16_kernel) and take the tensor-core path.
//   tc_linear_kernel         C = act(A W^T + bias + group_bias): persistent, warp-specialised
//                            (warp 0 TMA producer, warp 1 tcgen05.mma issuer, warps 2-5 epilogue),
//                            4-stage TMA->smem ring in the 128B-swizzled K-major UMMA layout,
//                            128 x BN fp32 accumulators double-buffered in TMEM so the epilogue of
//                            tile i overlaps the MMAs of tile i+1.  The epilogue reads TMEM with
//                            tcgen05.ld (one row per thread), fuses bias / per-group bias / ReLU,
//                            writes bf16 activations for the next layer and/or the max over each
//                            32-row patch (a lane-transpose reduction: 31 shuffles per 32 columns).
// The concat layer W.[g||f] is evaluated as W_g.g (tensor-core GEMM over groups, becomes the
// per-group bias) + W_f.f, so no (rows, 2E) tensor exists.  Activations are bf16 and are
// processed in chunks of groups sized to stay L2-resident between consecutive layers.
#include <cuda.h>
#include <cuda_bf16.h>
#include <stdlib.h>

#include "embed.cuh"

namespace p3tok {

// ------------------------------------------------------------------------------------------------ PTX
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a pipeline bug must surface as a trapped launch (reported through the C ABI), never as
// a hung GPU.  ~2 s at 2 GHz is orders of magnitude beyond any legitimate wait in these kernels.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000ll) {
      printf("p3tok: mbarrier wait timed out (block %d thread %d parity %u)\n", (int)blockIdx.x, (int)threadIdx.x, parity);
      __trap();
    }
  }
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T, kind::f16 (bf16 operands, fp32 accumulate)
__device__ __forceinline__ void tc_mma(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// issue only: the registers are valid after tc_ld_wait()
__device__ __forceinline__ void tc_ld32_issue(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tc_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// UMMA shared-memory descriptor: K-major operand, 128-byte swizzle, rows of 128 B, 8-row groups 1024 B apart
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3fffu);   // start address  [0,14)
  d |= (uint64_t)1 << 16;                    // leading byte offset (unused for swizzled K-major) [16,30)
  d |= (uint64_t)(1024 >> 4) << 32;          // stride byte offset [32,46)
  d |= (uint64_t)1 << 46;                    // descriptor version (Blackwell) [46,48)
  d |= (uint64_t)2 << 61;                    // SWIZZLE_128B [61,64)
  return d;
}

// ------------------------------------------------------------------------------------------------ GEMM
constexpr int TC_BM = 128, TC_BK = 64, TC_STAGES = 3, TC_MAX_BN = 256;
constexpr int TC_EPI_WARPS = 8;                      // two warps per TMEM lane quarter, alternating 64-column groups
constexpr int TC_THREADS = 64 + TC_EPI_WARPS * 32;   // warp 0 TMA, warp 1 MMA, warps 2..9 epilogue
constexpr int TC_A_STAGE = TC_BM * TC_BK * 2;        // 16 KB
constexpr int TC_B_STAGE = TC_MAX_BN * TC_BK * 2;    // 32 KB
constexpr int TC_STAGING = 32 * 128;                 // one 32-row x 64-column bf16 store box (4 KB, 128B-swizzled)
constexpr int TC_SMEM_PIPE = TC_STAGES * (TC_A_STAGE + TC_B_STAGE);
constexpr int TC_MAX_N = 2048;                      // bias vector staged in shared memory (8 KB)
constexpr int TC_SMEM = TC_SMEM_PIPE + TC_EPI_WARPS * 2 * TC_STAGING + TC_MAX_N * 4 + 256 + 1024;

struct TcParams {
  int M, N, K, BN;
  int num_m_tiles, num_n_tiles;
  const float* bias;       // [N] or null
  const float* gbias;      // [ceil(M/rows_per_group), N] or null
  int rows_per_group;
  int relu;
  int store_bf16;          // write bf16 [M,N] through tmC (TMA store)
  float* out_f32;          // [M,N] or null
  float* out_max;          // [ceil(M/32), N] max over each 32 consecutive rows, or null
  __nv_bfloat16* out_max_bf16;
  int max_relu;            // apply ReLU to the max (out_relu of the block)
};

// bias / per-group bias / ReLU on 32 accumulator columns starting at global column n0.  `sbias` is the
// bias vector staged in shared memory (zeros when the layer has none); all loads are issued before use.
__device__ __forceinline__ void epilogue_affine(float (&v)[32], const TcParams& p, const float* sbias, const float* gb, int n0) {
  const int nmax = p.N - 4;                                  // N % 8 == 0: clamped float4 loads stay in range
  float4 g4[8];
  if (gb) {
#pragma unroll
    for (int j = 0; j < 8; ++j) g4[j] = __ldg(reinterpret_cast<const float4*>(gb + min(n0 + 4 * j, nmax)));
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float4 b4 = *reinterpret_cast<const float4*>(sbias + min(n0 + 4 * j, nmax));
    v[4 * j] += b4.x; v[4 * j + 1] += b4.y; v[4 * j + 2] += b4.z; v[4 * j + 3] += b4.w;
  }
  if (gb) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      v[4 * j] += g4[j].x; v[4 * j + 1] += g4[j].y; v[4 * j + 2] += g4[j].z; v[4 * j + 3] += g4[j].w;
    }
  }
  if (p.relu) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
  }
}

// max over the warp's 32 rows (lane = row): lane-transpose reduction, lane l returns column l
__device__ __forceinline__ float warp_rows_max(float (&v)[32], int lane) {
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

__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
  __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&t);
}

__global__ void __launch_bounds__(TC_THREADS, 1)
tc_linear_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const __grid_constant__ CUtensorMap tmC, const TcParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;
  uint8_t* sB = smem + TC_STAGES * TC_A_STAGE;
  uint8_t* sC = smem + TC_SMEM_PIPE;                  // per-epilogue-warp store staging, 2 x 4 KB each
  float* sbias = reinterpret_cast<float*>(sC + TC_EPI_WARPS * 2 * TC_STAGING);
  uint64_t* full = reinterpret_cast<uint64_t*>(sC + TC_EPI_WARPS * 2 * TC_STAGING + TC_MAX_N * 4);
  uint64_t* empty = full + TC_STAGES;
  uint64_t* tfull = empty + TC_STAGES;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_tiles = p.num_m_tiles * p.num_n_tiles;
  const int num_kb = (p.K + TC_BK - 1) / TC_BK;

  if (threadIdx.x == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmA)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmB)) : "memory");
    if (p.store_bf16) asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmC)) : "memory");
    for (int s = 0; s < TC_STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&tfull[b], 1);
      mbar_init(&tempty[b], TC_EPI_WARPS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = threadIdx.x; i < p.N; i += TC_THREADS) sbias[i] = p.bias ? p.bias[i] : 0.f;
  if (warp == 1) {   // TMEM: 512 columns = two 128 x BN fp32 accumulators
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {   // ---------------- TMA producer
      int stage = 0;
      uint32_t phase = 0;
      const uint32_t tx = TC_A_STAGE + (uint32_t)p.BN * TC_BK * 2;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int mt = tile / p.num_n_tiles, nt = tile - mt * p.num_n_tiles;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty[stage], phase ^ 1);
          mbar_expect_tx(&full[stage], tx);
          tma_load_2d(sA + stage * TC_A_STAGE, &tmA, &full[stage], kb * TC_BK, mt * TC_BM);
          tma_load_2d(sB + stage * TC_B_STAGE, &tmB, &full[stage], kb * TC_BK, nt * p.BN);
          if (++stage == TC_STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {   // ---------------- MMA issuer
      // instruction descriptor: D=f32, A=B=bf16, both K-major, N = BN, M = 128
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.BN >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
        const int buf = it & 1;
        mbar_wait(&tempty[buf], ((uint32_t)(it >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(buf * p.BN);
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint32_t a0 = smem_u32(sA + stage * TC_A_STAGE), b0 = smem_u32(sB + stage * TC_B_STAGE);
#pragma unroll
          for (int k = 0; k < TC_BK / 16; ++k)
            tc_mma(tmem_d, umma_desc_sw128(a0 + k * 32), umma_desc_sw128(b0 + k * 32), idesc, (uint32_t)((kb | k) != 0));
          tc_commit(&empty[stage]);   // frees the smem stage once these MMAs have read it
          if (++stage == TC_STAGES) { stage = 0; phase ^= 1; }
        }
        tc_commit(&tfull[buf]);       // accumulator complete
      }
    }
  } else {             // ---------------- epilogue warps: TMEM lane quarter q, column-group parity h
    const int ew = warp - 2;
    const int q = warp & 3, h = ew >> 2;
    uint8_t* stg = sC + ew * 2 * TC_STAGING;
    int sbuf = 0;
    int it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const int mt = tile / p.num_n_tiles, nt = tile - mt * p.num_n_tiles;
      const int buf = it & 1;
      mbar_wait(&tfull[buf], (uint32_t)(it >> 1) & 1);
      tc_fence_after();
      const int row0 = mt * TC_BM + q * 32;
      const int row = row0 + lane;
      const bool row_ok = row < p.M;
      const float* gb = (p.gbias && row_ok) ? p.gbias + (size_t)(row / p.rows_per_group) * p.N : nullptr;
      for (int gi = h; gi * 64 < p.BN; gi += 2) {
        const int n0 = nt * p.BN + gi * 64;
        if (n0 >= p.N || row0 >= p.M) break;     // warp-uniform
        float v0[32], v1[32];
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * p.BN + gi * 64);
        tc_ld32_issue(taddr, v0);
        tc_ld32_issue(taddr + 32, v1);
        tc_ld_wait();
        epilogue_affine(v0, p, sbias, gb, n0);
        epilogue_affine(v1, p, sbias, gb, n0 + 32);
        if (p.store_bf16) {
          // stage the 32 x 64 bf16 box in the 128B-swizzled layout the store tensor map expects, then one
          // TMA store (clips rows >= M and columns >= N); double-buffered per warp
          uint8_t* sb = stg + sbuf * TC_STAGING;
          if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
          __syncwarp();
          const uint32_t rbase = smem_u32(sb) + lane * 128;
#pragma unroll
          for (int pc = 0; pc < 4; ++pc) {
            const uint32_t a0 = rbase + (((uint32_t)pc ^ (lane & 7)) << 4);
            const uint32_t a1 = rbase + (((uint32_t)(pc + 4) ^ (lane & 7)) << 4);
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a0), "r"(pack_bf16x2(v0[pc * 8], v0[pc * 8 + 1])),
                         "r"(pack_bf16x2(v0[pc * 8 + 2], v0[pc * 8 + 3])), "r"(pack_bf16x2(v0[pc * 8 + 4], v0[pc * 8 + 5])),
                         "r"(pack_bf16x2(v0[pc * 8 + 6], v0[pc * 8 + 7]))
                         : "memory");
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a1), "r"(pack_bf16x2(v1[pc * 8], v1[pc * 8 + 1])),
                         "r"(pack_bf16x2(v1[pc * 8 + 2], v1[pc * 8 + 3])), "r"(pack_bf16x2(v1[pc * 8 + 4], v1[pc * 8 + 5])),
                         "r"(pack_bf16x2(v1[pc * 8 + 6], v1[pc * 8 + 7]))
                         : "memory");
          }
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          __syncwarp();
          if (lane == 0) {
            asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                             reinterpret_cast<uint64_t>(&tmC)),
                         "r"(smem_u32(sb)), "r"(n0), "r"(row0)
                         : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          }
          sbuf ^= 1;
        }
        if (p.out_f32 && row_ok) {
          float* o = p.out_f32 + (size_t)row * p.N + n0;
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            if (n0 + j < p.N) *reinterpret_cast<float4*>(o + j) = make_float4(v0[j], v0[j + 1], v0[j + 2], v0[j + 3]);
            if (n0 + 32 + j < p.N) *reinterpret_cast<float4*>(o + 32 + j) = make_float4(v1[j], v1[j + 1], v1[j + 2], v1[j + 3]);
          }
        }
        if (p.out_max || p.out_max_bf16) {
          if (!row_ok) {
#pragma unroll
            for (int j = 0; j < 32; ++j) { v0[j] = -3.0e38f; v1[j] = -3.0e38f; }
          }
          float m0 = warp_rows_max(v0, lane), m1 = warp_rows_max(v1, lane);
          if (p.max_relu) { m0 = fmaxf(m0, 0.f); m1 = fmaxf(m1, 0.f); }
          const size_t grow = (size_t)(row0 >> 5) * p.N;
          const int na = n0 + lane, nb = n0 + 32 + lane;
          if (p.out_max) {
            if (na < p.N) p.out_max[grow + na] = m0;
            if (nb < p.N) p.out_max[grow + nb] = m1;
          }
          if (p.out_max_bf16) {
            if (na < p.N) p.out_max_bf16[grow + na] = __float2bfloat16_rn(m0);
            if (nb < p.N) p.out_max_bf16[grow + nb] = __float2bfloat16_rn(m1);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[buf]);
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // all stores of this warp are complete
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
  }
}

// ------------------------------------------------------------------------------------------------ host: TMA maps
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// bf16 row-major [rows, cols] (row pitch = cols*2 bytes), box = 64 cols x box_rows, 128B swizzle, OOB -> 0
static int make_map(CUtensorMap* m, const void* base, int64_t rows, int64_t cols, int box_rows) {
  P3_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0, P3TOK_ERR_UNSUPPORTED, "tensor map: base must be 16-byte aligned");
  EncodeTiledFn fn = encode_fn();
  P3_REQUIRE(fn, P3TOK_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)cols * 2};
  cuuint32_t box[2] = {(cuuint32_t)TC_BK, (cuuint32_t)box_rows};
  cuuint32_t es[2] = {1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, es,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  P3_REQUIRE(r == CUDA_SUCCESS, P3TOK_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d) rows=%lld cols=%lld box_rows=%d", (int)r,
             (long long)rows, (long long)cols, box_rows);
  return P3TOK_OK;
}

static int num_sms() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

static int pick_bn(int N) {
  // widest tile <= 256 (multiple of 64: the epilogue works in 64-column store boxes) with the least padding
  int best = 64, best_waste = 1 << 30;
  for (int bn = 256; bn >= 64; bn -= 64) {
    const int tiles = (N + bn - 1) / bn;
    const int waste = tiles * bn - N;
    if (waste < best_waste) { best_waste = waste; best = bn; }
  }
  return best;
}

// C = act(A[M,K] W[N,K]^T + bias + gbias), bf16 operands.  K % 8 == 0, N % 8 == 0.
static int tc_linear(const __nv_bfloat16* A, int64_t M, int K, const __nv_bfloat16* W, int N, const float* bias,
                     const float* gbias, int rows_per_group, int relu, __nv_bfloat16* out_bf16, float* out_f32,
                     float* out_max, __nv_bfloat16* out_max_bf16, int max_relu, cudaStream_t s) {
  P3_REQUIRE(K % 8 == 0 && N % 8 == 0, P3TOK_ERR_UNSUPPORTED, "tc_linear: K=%d and N=%d must be multiples of 8", K, N);
  P3_REQUIRE(N <= TC_MAX_N, P3TOK_ERR_UNSUPPORTED, "tc_linear: N=%d > %d", N, TC_MAX_N);
  P3_REQUIRE(M < (1ll << 31) - 256, P3TOK_ERR_UNSUPPORTED, "tc_linear: too many rows");
  if (M == 0) return P3TOK_OK;
  TcParams p;
  p.M = (int)M; p.N = N; p.K = K; p.BN = pick_bn(N);
  p.num_m_tiles = (int)((M + TC_BM - 1) / TC_BM);
  p.num_n_tiles = (N + p.BN - 1) / p.BN;
  p.bias = bias; p.gbias = gbias; p.rows_per_group = rows_per_group > 0 ? rows_per_group : 1; p.relu = relu;
  p.store_bf16 = out_bf16 != nullptr; p.out_f32 = out_f32; p.out_max = out_max; p.out_max_bf16 = out_max_bf16; p.max_relu = max_relu;
  CUtensorMap ta, tb, tc;
  int rc = make_map(&ta, A, M, K, TC_BM);
  if (rc) return rc;
  rc = make_map(&tb, W, N, K, p.BN);
  if (rc) return rc;
  if (out_bf16) {
    rc = make_map(&tc, out_bf16, M, N, 32);     // store boxes: 64 columns x 32 rows
    if (rc) return rc;
  } else {
    tc = ta;                                     // unused by the kernel
  }
  static thread_local bool configured[32] = {false};
  int dev = 0;
  P3_CUDA(cudaGetDevice(&dev));
  if (dev < 32 && !configured[dev]) {
    P3_CUDA(cudaFuncSetAttribute(tc_linear_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM));
    configured[dev] = true;
  }
  const int tiles = p.num_m_tiles * p.num_n_tiles;
  const int grid = tiles < num_sms() ? tiles : num_sms();
  tc_linear_kernel<<<grid, TC_THREADS, TC_SMEM, s>>>(ta, tb, tc, p);
  P3_LAUNCH_CHECK("tc_linear_kernel");
  return P3TOK_OK;
}

// ------------------------------------------------------------------------------------------------ first layer
// Narrow input: one CTA = 32 rows (one k=32 patch when aligned), 128 threads x 2 output channels per pass.
template <typename IdxT, int CP>   // CP = padded input width (8 or 16)
__global__ void __launch_bounds__(128)
rows_first_layer_kernel(p3tok_rows R, int64_t g_begin, int64_t nrows, const __nv_bfloat16* __restrict__ W,
                        const float* __restrict__ bias, int cin, int nout, int relu, __nv_bfloat16* __restrict__ out) {
  __shared__ __align__(16) float xin[32][CP];
  const int64_t r0 = (int64_t)blockIdx.x * 32;
  const IdxT* knn = reinterpret_cast<const IdxT*>(R.knn_idx);
  for (int e = threadIdx.x; e < 32 * CP; e += 128) {
    const int rr = e / CP, c = e % CP;
    const int64_t r = r0 + rr;
    float v = 0.f;          // columns >= cin and rows >= nrows stay zero
    if (r < nrows && c < cin) {
      if (R.kind == 2) {
        v = R.x[(g_begin * R.k + r) * cin + c];
      } else {
        const int n = (int)(r % R.k);
        const int64_t bj = g_begin + r / R.k;
        const int64_t b = bj / R.G;
        const int64_t g = R.perm ? R.perm[bj] : (bj - b * R.G);
        const int64_t ni = (int64_t)knn[(b * R.G + g) * R.k + n];
        if (R.kind == 0) {
          const float* crow = R.x + (b * R.N + R.ctr_idx[b * R.G + g]) * R.C;
          v = c < R.C ? __fsub_rn(R.x[(b * R.N + ni) * R.C + c], crow[c]) : crow[c - R.C];
        } else {
          v = c < 3 ? R.x[(b * R.N + ni) * 3 + c] : R.feats[(b * R.N + ni) * R.D + (c - 3)];
        }
      }
    }
    xin[rr][c] = v;
  }
  __syncthreads();
  for (int n0 = threadIdx.x * 2; n0 < nout; n0 += 256) {
    float w0[CP], w1[CP];
#pragma unroll
    for (int c = 0; c < CP; ++c) {
      w0[c] = c < cin ? __bfloat162float(W[(size_t)n0 * cin + c]) : 0.f;
      w1[c] = (c < cin && n0 + 1 < nout) ? __bfloat162float(W[(size_t)(n0 + 1) * cin + c]) : 0.f;
    }
    const float b0 = bias ? bias[n0] : 0.f, b1 = (bias && n0 + 1 < nout) ? bias[n0 + 1] : 0.f;
    for (int rr = 0; rr < 32; ++rr) {
      const int64_t r = r0 + rr;
      if (r >= nrows) break;
      float a0 = b0, a1 = b1;
#pragma unroll
      for (int c4 = 0; c4 < CP; c4 += 4) {
        const float4 xv = *reinterpret_cast<const float4*>(&xin[rr][c4]);   // warp-wide broadcast
        a0 = fmaf(w0[c4], xv.x, a0); a1 = fmaf(w1[c4], xv.x, a1);
        a0 = fmaf(w0[c4 + 1], xv.y, a0); a1 = fmaf(w1[c4 + 1], xv.y, a1);
        a0 = fmaf(w0[c4 + 2], xv.z, a0); a1 = fmaf(w1[c4 + 2], xv.z, a1);
        a0 = fmaf(w0[c4 + 3], xv.w, a0); a1 = fmaf(w1[c4 + 3], xv.w, a1);
      }
      if (relu) { a0 = fmaxf(a0, 0.f); a1 = fmaxf(a1, 0.f); }
      *reinterpret_cast<__nv_bfloat162*>(out + r * nout + n0) = __floats2bfloat162_rn(a0, a1);
    }
  }
}

// Wide input: gather rows to bf16 [nrows, kpad], zero padded
template <typename IdxT>
__global__ void rows_gather_bf16_kernel(p3tok_rows R, int64_t g_begin, int64_t nrows, int cin, int kpad,
                                        __nv_bfloat16* __restrict__ out) {
  const IdxT* knn = reinterpret_cast<const IdxT*>(R.knn_idx);
  const int64_t total = nrows * kpad;
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(e % kpad);
    const int64_t r = e / kpad;
    float v = 0.f;
    if (c < cin) {
      if (R.kind == 2) {
        v = R.x[(g_begin * R.k + r) * cin + c];
      } else {
        const int n = (int)(r % R.k);
        const int64_t bj = g_begin + r / R.k;
        const int64_t b = bj / R.G;
        const int64_t g = R.perm ? R.perm[bj] : (bj - b * R.G);
        const int64_t ni = (int64_t)knn[(b * R.G + g) * R.k + n];
        if (R.kind == 0) {
          const float* crow = R.x + (b * R.N + R.ctr_idx[b * R.G + g]) * R.C;
          v = c < R.C ? __fsub_rn(R.x[(b * R.N + ni) * R.C + c], crow[c]) : crow[c - R.C];
        } else {
          v = c < 3 ? R.x[(b * R.N + ni) * 3 + c] : R.feats[(b * R.N + ni) * R.D + (c - 3)];
        }
      }
    }
    out[e] = __float2bfloat16_rn(v);
  }
}

__global__ void pad_weight_kernel(const __nv_bfloat16* __restrict__ W, int N, int K, int kpad, __nv_bfloat16* __restrict__ out) {
  const int total = N * kpad;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
    const int c = e % kpad, n = e / kpad;
    out[e] = c < K ? W[(size_t)n * K + c] : __float2bfloat16_rn(0.f);
  }
}

// max over `parts` consecutive partial rows (k = 32*parts): in [ngroups*parts, C] -> out [ngroups, C]
__global__ void partial_max_kernel(const float* __restrict__ in, int64_t ngroups, int parts, int C, float* out_f32,
                                   __nv_bfloat16* out_bf16) {
  const int64_t total = ngroups * C;
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t g = e / C;
    const int c = (int)(e - g * C);
    float m = in[(g * parts) * C + c];
    for (int q = 1; q < parts; ++q) m = fmaxf(m, in[(g * parts + q) * C + c]);
    if (out_f32) out_f32[e] = m;
    if (out_bf16) out_bf16[e] = __float2bfloat16_rn(m);
  }
}

// generic group max for k not a multiple of 32: in fp32 [ngroups*k, C]
__global__ void group_max_f32_kernel(const float* __restrict__ in, int64_t ngroups, int k, int C, int relu, float* out_f32,
                                     __nv_bfloat16* out_bf16) {
  const int64_t total = ngroups * C;
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t g = e / C;
    const int c = (int)(e - g * C);
    const float* q = in + g * k * C + c;
    float m = q[0];
    for (int r = 1; r < k; ++r) m = fmaxf(m, q[(int64_t)r * C]);
    if (relu) m = fmaxf(m, 0.f);
    if (out_f32) out_f32[e] = m;
    if (out_bf16) out_bf16[e] = __float2bfloat16_rn(m);
  }
}

__global__ void group_max_bf16_kernel(const __nv_bfloat16* __restrict__ in, int64_t ngroups, int k, int C,
                                      __nv_bfloat16* __restrict__ out) {
  const int64_t total = ngroups * C;
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t g = e / C;
    const int c = (int)(e - g * C);
    const __nv_bfloat16* q = in + g * k * C + c;
    float m = __bfloat162float(q[0]);
    for (int r = 1; r < k; ++r) m = fmaxf(m, __bfloat162float(q[(int64_t)r * C]));
    out[e] = __float2bfloat16_rn(m);
  }
}

static inline unsigned grid_1d(int64_t total, int threads) {
  int64_t b = (total + threads - 1) / threads;
  const int64_t cap = 148 * 16;
  return (unsigned)(b < 1 ? 1 : (b > cap ? cap : b));
}

// ------------------------------------------------------------------------------------------------ orchestration
// Chunking: one chunk of rows goes through all layers before the next starts.  Measured on B200 (C2,
// 524288 rows): per-launch fixed cost (prologue, pipeline fill, last-tile epilogue) outweighs keeping a
// chunk's activations L2-resident - 2-wave chunks 1.91 ms, 16-wave chunks 1.35 ms - so chunks are as
// large as a bounded workspace allows (default 2^20 rows: <= 1.6 GB per bf16 activation buffer at 768
// columns).  P3TOK_CHUNK_ROWS overrides for experiments.
static int64_t chunk_groups(int64_t ngroups, int64_t k) {
  static int64_t rows_target = 0;
  if (!rows_target) {
    const char* e = getenv("P3TOK_CHUNK_ROWS");
    rows_target = e ? atoll(e) : (1ll << 20);
    if (rows_target < 128) rows_target = 128;
  }
  int64_t cg = rows_target / k;
  if (cg < 1) cg = 1;
  return ngroups < cg ? ngroups : cg;
}

struct BfLayout {
  int64_t cg, rows, wmax, F, kpad0;
  int64_t off_act0, off_act1, off_gmax_f32, off_gmax_bf16, off_gbias, off_scratch_f32, off_wpad, total;
};

static BfLayout bf_layout(const p3tok_mlp* m, int64_t ngroups, int64_t k) {
  BfLayout L;
  L.cg = chunk_groups(ngroups, k);
  L.rows = L.cg * k;
  L.F = m->pre_dim[m->n_pre - 1];
  L.kpad0 = m->cin <= 16 ? 0 : align_up(m->cin, 8);
  int64_t w = L.kpad0;
  for (int i = 0; i < m->n_pre; ++i) w = w > m->pre_dim[i] ? w : m->pre_dim[i];
  w = w > m->mid_dim ? w : m->mid_dim;
  L.wmax = w;
  const int64_t wide = L.F > m->out_dim ? L.F : m->out_dim;
  int64_t o = 0;
  L.off_act0 = o; o += align_up(L.rows * w * 2, 1024);
  L.off_act1 = o; o += align_up(L.rows * w * 2, 1024);
  L.off_gmax_f32 = o; o += align_up((L.rows / 32 + 1) * wide * 4, 1024);
  L.off_gmax_bf16 = o; o += align_up((L.cg + 1) * L.F * 2, 1024);
  L.off_gbias = o; o += align_up(L.cg * m->mid_dim * 4, 1024);
  // fp32 scratch for the unfused max (k not a multiple of 32): one [rows, max(F,out)] matrix
  L.off_scratch_f32 = o; o += (k % 32 == 0) ? 0 : align_up(L.rows * wide * 4, 1024);
  L.off_wpad = o; o += L.kpad0 ? align_up((int64_t)m->pre_dim[0] * L.kpad0 * 2, 1024) : 0;
  L.total = o + 1024;
  return L;
}

int64_t patch_embed_bf16_workspace(const p3tok_mlp* m, int64_t ngroups, int64_t k) { return bf_layout(m, ngroups, k).total; }

int patch_embed_bf16(const p3tok_rows* R, const p3tok_mlp* m, void* ws, int64_t ws_bytes, float* tokens, cudaStream_t s) {
  const int64_t ngroups = R->B * R->G, k = R->k;
  const BfLayout L = bf_layout(m, ngroups, k);
  P3_REQUIRE(ws_bytes >= L.total, P3TOK_ERR_WORKSPACE, "patch_embed(bf16): workspace %lld < %lld bytes", (long long)ws_bytes,
             (long long)L.total);
  for (int i = 0; i < m->n_pre; ++i)
    P3_REQUIRE(m->pre_dim[i] % 8 == 0, P3TOK_ERR_UNSUPPORTED, "patch_embed(bf16): layer widths must be multiples of 8");
  P3_REQUIRE(m->mid_dim % 8 == 0 && m->out_dim % 8 == 0, P3TOK_ERR_UNSUPPORTED,
             "patch_embed(bf16): layer widths must be multiples of 8");
  char* base = reinterpret_cast<char*>((reinterpret_cast<uintptr_t>(ws) + 1023) / 1024 * 1024);
  __nv_bfloat16* act[2] = {reinterpret_cast<__nv_bfloat16*>(base + L.off_act0), reinterpret_cast<__nv_bfloat16*>(base + L.off_act1)};
  float* gmax_f32 = reinterpret_cast<float*>(base + L.off_gmax_f32);
  __nv_bfloat16* gmax_bf16 = reinterpret_cast<__nv_bfloat16*>(base + L.off_gmax_bf16);
  float* gbias = reinterpret_cast<float*>(base + L.off_gbias);
  float* scratch = reinterpret_cast<float*>(base + L.off_scratch_f32);
  __nv_bfloat16* wpad = reinterpret_cast<__nv_bfloat16*>(base + L.off_wpad);
  const bool fused_max = (k % 32 == 0);
  const int parts = fused_max ? (int)(k / 32) : 0;
  const bool i64 = (R->kind == 2) || R->idx_dtype == P3TOK_I64;

  if (L.kpad0) {
    pad_weight_kernel<<<grid_1d((int64_t)m->pre_dim[0] * L.kpad0, 256), 256, 0, s>>>(
        (const __nv_bfloat16*)m->w_pre[0], m->pre_dim[0], m->cin, (int)L.kpad0, wpad);
    P3_LAUNCH_CHECK("pad_weight_kernel");
  }

  for (int64_t g0 = 0; g0 < ngroups; g0 += L.cg) {
    const int64_t gc = (ngroups - g0) < L.cg ? (ngroups - g0) : L.cg;
    const int64_t rows = gc * k;
    int cur = 0;
    int rc;
    // ---- first layer: CUDA cores for narrow inputs, otherwise gather to bf16 rows for the tensor cores
    int first_tc = 0;
    int kin;
    if (!L.kpad0) {
      const unsigned blocks = (unsigned)((rows + 31) / 32);
#define P3_L1(IDX, CPV)                                                                                        \
  rows_first_layer_kernel<IDX, CPV><<<blocks, 128, 0, s>>>(*R, g0, rows, (const __nv_bfloat16*)m->w_pre[0],    \
                                                           m->b_pre[0], m->cin, m->pre_dim[0], m->pre_relu[0], act[cur])
      if (m->cin <= 8) { if (i64) P3_L1(int64_t, 8); else P3_L1(int32_t, 8); }
      else             { if (i64) P3_L1(int64_t, 16); else P3_L1(int32_t, 16); }
#undef P3_L1
      P3_LAUNCH_CHECK("rows_first_layer_kernel");
      first_tc = 1;
      kin = m->pre_dim[0];
    } else {
      if (i64)
        rows_gather_bf16_kernel<int64_t><<<grid_1d(rows * L.kpad0, 256), 256, 0, s>>>(*R, g0, rows, m->cin, (int)L.kpad0, act[cur]);
      else
        rows_gather_bf16_kernel<int32_t><<<grid_1d(rows * L.kpad0, 256), 256, 0, s>>>(*R, g0, rows, m->cin, (int)L.kpad0, act[cur]);
      P3_LAUNCH_CHECK("rows_gather_bf16_kernel");
      kin = (int)L.kpad0;
    }
    // ---- remaining per-point layers on the tensor cores; the last one also emits the patch max
    bool have_gmax = false;
    for (int i = first_tc; i < m->n_pre; ++i) {
      const bool last = (i == m->n_pre - 1);
      const __nv_bfloat16* Wi = (i == 0 && L.kpad0) ? wpad : (const __nv_bfloat16*)m->w_pre[i];
      if (last && fused_max) {
        rc = tc_linear(act[cur], rows, kin, Wi, m->pre_dim[i], m->b_pre[i], nullptr, 1, m->pre_relu[i], act[cur ^ 1], nullptr,
                       parts == 1 ? nullptr : gmax_f32, parts == 1 ? gmax_bf16 : nullptr, 0, s);
        if (rc) return rc;
        if (parts > 1) {
          partial_max_kernel<<<grid_1d(gc * L.F, 256), 256, 0, s>>>(gmax_f32, gc, parts, (int)L.F, nullptr, gmax_bf16);
          P3_LAUNCH_CHECK("partial_max_kernel");
        }
        have_gmax = true;
      } else if (last) {
        rc = tc_linear(act[cur], rows, kin, Wi, m->pre_dim[i], m->b_pre[i], nullptr, 1, m->pre_relu[i], act[cur ^ 1], scratch,
                       nullptr, nullptr, 0, s);
        if (rc) return rc;
        group_max_f32_kernel<<<grid_1d(gc * L.F, 256), 256, 0, s>>>(scratch, gc, (int)k, (int)L.F, 0, nullptr, gmax_bf16);
        P3_LAUNCH_CHECK("group_max_f32_kernel");
        have_gmax = true;
      } else {
        rc = tc_linear(act[cur], rows, kin, Wi, m->pre_dim[i], m->b_pre[i], nullptr, 1, m->pre_relu[i], act[cur ^ 1], nullptr,
                       nullptr, nullptr, 0, s);
        if (rc) return rc;
      }
      cur ^= 1;
      kin = m->pre_dim[i];
    }
    if (!have_gmax) {
      // the block's only per-point layer ran on CUDA cores (P3Embed stage 0): reduce its bf16 output
      group_max_bf16_kernel<<<grid_1d(gc * L.F, 256), 256, 0, s>>>(act[cur], gc, (int)k, (int)L.F, gmax_bf16);
      P3_LAUNCH_CHECK("group_max_bf16_kernel");
    }
    // ---- concat layer: per-group half (becomes a bias), then the per-point half
    rc = tc_linear(gmax_bf16, gc, (int)L.F, (const __nv_bfloat16*)m->w_mid_g, m->mid_dim, m->b_mid, nullptr, 1, 0, nullptr, gbias,
                   nullptr, nullptr, 0, s);
    if (rc) return rc;
    rc = tc_linear(act[cur], rows, (int)L.F, (const __nv_bfloat16*)m->w_mid_f, m->mid_dim, nullptr, gbias, (int)k, 1, act[cur ^ 1],
                   nullptr, nullptr, nullptr, 0, s);
    if (rc) return rc;
    cur ^= 1;
    // ---- output layer + max over the patch
    float* tok = tokens + g0 * m->out_dim;
    if (fused_max) {
      rc = tc_linear(act[cur], rows, m->mid_dim, (const __nv_bfloat16*)m->w_out, m->out_dim, m->b_out, nullptr, 1, 0, nullptr, nullptr,
                     parts == 1 ? tok : gmax_f32, nullptr, parts == 1 ? m->out_relu : 0, s);
      if (rc) return rc;
      if (parts > 1) {
        // ReLU commutes with max: apply it after combining the partial maxima
        group_max_f32_kernel<<<grid_1d(gc * m->out_dim, 256), 256, 0, s>>>(gmax_f32, gc, parts, m->out_dim, m->out_relu, tok, nullptr);
        P3_LAUNCH_CHECK("group_max_f32_kernel");
      }
    } else {
      rc = tc_linear(act[cur], rows, m->mid_dim, (const __nv_bfloat16*)m->w_out, m->out_dim, m->b_out, nullptr, 1, 0, nullptr, scratch,
                     nullptr, nullptr, 0, s);
      if (rc) return rc;
      group_max_f32_kernel<<<grid_1d(gc * m->out_dim, 256), 256, 0, s>>>(scratch, gc, (int)k, m->out_dim, m->out_relu, tok, nullptr);
      P3_LAUNCH_CHECK("group_max_f32_kernel");
    }
  }
  return P3TOK_OK;
}

}  // namespace p3tok

using namespace p3tok;

extern "C" int p3tok_linear_bf16(const void* A, int64_t M, int64_t K, const void* W, int64_t N, const float* bias,
                                 const float* gbias, int64_t rows_per_group, int relu, void* out_bf16, float* out_f32,
                                 float* out_max32, void* stream) {
  P3_REQUIRE(M >= 0 && K > 0 && N > 0 && K < (1 << 24) && N < (1 << 24), P3TOK_ERR_INVALID, "linear_bf16: bad shape");
  if (M == 0) return P3TOK_OK;
  P3_REQUIRE(A && W && (out_bf16 || out_f32 || out_max32), P3TOK_ERR_INVALID, "linear_bf16: null pointer");
  P3_REQUIRE((reinterpret_cast<uintptr_t>(A) & 15) == 0 && (reinterpret_cast<uintptr_t>(W) & 15) == 0, P3TOK_ERR_UNSUPPORTED,
             "linear_bf16: operands must be 16-byte aligned");
  return tc_linear((const __nv_bfloat16*)A, M, (int)K, (const __nv_bfloat16*)W, (int)N, bias, gbias, (int)rows_per_group, relu,
                   (__nv_bfloat16*)out_bf16, out_f32, out_max32, nullptr, 0, as_stream(stream));
}

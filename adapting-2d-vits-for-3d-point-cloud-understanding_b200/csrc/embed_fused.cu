// embed_fused.cu - two consecutive layers of the patch-embedding MLP in ONE tcgen05 kernel.
//
//   out = W_b . relu(W_a . A0 + bias_a [+ group_bias]) + bias_b          (then bf16 store and/or 32-row max)
//
// used twice per patch-embedding block (reference src/models/apf.py:129-143, src/models/pix4point.py:147-156):
//   "pre"  pair  h1 -> 256->512 (ReLU) -> 512->E        emits f (bf16) and the per-patch max g
//   "post" pair  f  -> E->2E (+W_g.g as group bias, ReLU) -> 2E->E   emits the patch tokens (max over the patch)
// The hidden activation (rows x 512 / rows x 768, the widest tensors of the block) never leaves the SM:
// per 128-row tile the first GEMM is evaluated in 64-column chunks into a small TMEM accumulator, the
// epilogue warps turn each chunk into a bf16 K-major shared-memory operand (bias, ReLU), and the second GEMM
// consumes it at once, accumulating the tile's full output row block in TMEM (<= 384 columns).  Measured
// motivation (profiles/): as separate GEMMs these layers are bound by their store epilogues and
// by streaming the activation tile from L2/HBM, not by the tensor pipe.
//
// Shared memory per CTA: A0 tile (K0/64 x 16 KB, resident for the tile) | 2 chunk operand buffers (16 KB) |
// ring RA of 8 KB weight boxes [64 n x 64 k] of W_a | ring RB of boxes [N2/4 n x 64 k] of W_b | biases.
// TMEM: columns [0,N2) output accumulator, [384,448) and [448,512) the two chunk accumulators.
// Warps: 0 = TMA producer (A0 + RA), 1 = MMA issuer, 2..9 = epilogue (quarter q = warp%4 of the 128 rows,
// half h of each 64-column chunk / alternating 64-column groups of the output), 10 = TMA producer (RB).
#include <vector>

#include "tc_common.cuh"

namespace p3tok {

constexpr int FU_EPI_WARPS = 16;              // 4 per row quarter: each converts a 16-column slice of every chunk
constexpr int FU_THREADS = (3 + FU_EPI_WARPS) * 32;
constexpr int FU_CHUNK = 64;                 // hidden columns per chunk
constexpr int FU_ACC2_COL = 384;             // TMEM column of chunk accumulator 0
constexpr int FU_RA_BOX = 64 * 128;          // 8 KB: 64 rows of W_a x 64 k
constexpr int FU_MAX_RA = 16, FU_MAX_RB = 8;
constexpr int FU_SMEM = 227 * 1024;

// 16 accumulator columns of this warp's 32 TMEM lanes (lane = row)
__device__ __forceinline__ void tc_ld16_issue(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

struct FusedParams {
  int M, K0, N1, N2;
  int num_m_tiles;
  int ra_slots, rb_slots, rb_box;   // ring depths; rb_box = (N2/4) * 128 bytes
  const float* bias_a;              // [N1] or null
  const float* gbias;               // [M / rows_per_group, N1] or null; rows_per_group % 32 == 0
  int rows_per_group;
  const float* bias_b;              // [N2] or null
  int store_out;                    // bf16 [M,N2] through tmC
  float* out_max;                   // [M/32, N2] or null
  __nv_bfloat16* out_max_bf16;      // same, bf16, or null
  int max_relu;
  unsigned long long* trace;        // debug (P3TOK_TC_TRACE=1): CTA 0's MMA-warp timeline, 8 stamps per chunk
};
__device__ __forceinline__ void fu_trace(const FusedParams& p, int it, int j, int slot, long long v) {
  if (p.trace && blockIdx.x == 0 && it < 4 && j < 16) p.trace[((size_t)it * 16 + j) * 8 + slot] = (unsigned long long)v;
}

__global__ void __launch_bounds__(FU_THREADS, 1)
tc_fused_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmWa,
                const __grid_constant__ CUtensorMap tmWb, const __grid_constant__ CUtensorMap tmC, const FusedParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int KB0 = p.K0 / 64, NC = p.N1 / FU_CHUNK, NQ = 4;
  uint8_t* sA0 = smem;                                   // KB0 x 16 KB
  uint8_t* sCH = sA0 + KB0 * 16384;                      // 2 x 16 KB (also the epilogue's store staging)
  uint8_t* sRA = sCH + 2 * 16384;                        // ra_slots x 8 KB
  uint8_t* sRB = sRA + p.ra_slots * FU_RA_BOX;           // rb_slots x rb_box
  float* sba = reinterpret_cast<float*>(sRB + p.rb_slots * p.rb_box);   // N1 floats
  float* sbb = sba + p.N1;                               // N2 floats
  float* sgb = sbb + p.N2;                               // 8 warps x 64 floats (group-bias slices)
  uint64_t* bars = reinterpret_cast<uint64_t*>((reinterpret_cast<uintptr_t>(sgb + 8 * 64) + 7) & ~(uintptr_t)7);
  uint64_t* a0_full = bars;            // 1
  uint64_t* a0_empty = bars + 1;       // 1
  uint64_t* acc3_full = bars + 2;      // 1
  uint64_t* acc3_empty = bars + 3;     // 1
  uint64_t* acc2_full = bars + 4;      // 2
  uint64_t* acc2_empty = bars + 6;     // 2
  uint64_t* ch_full = bars + 8;        // 2
  uint64_t* ch_empty = bars + 10;      // 2
  uint64_t* ra_full = bars + 12;       // 16
  uint64_t* ra_empty = bars + 28;      // 16
  uint64_t* rb_full = bars + 44;       // 8
  uint64_t* rb_empty = bars + 52;      // 8
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 60);

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmA)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmWa)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmWb)) : "memory");
    mbar_init(a0_full, 1);
    mbar_init(a0_empty, 1);
    mbar_init(acc3_full, 1);
    mbar_init(acc3_empty, FU_EPI_WARPS);
    for (int b = 0; b < 2; ++b) {
      mbar_init(&acc2_full[b], 1);
      mbar_init(&acc2_empty[b], FU_EPI_WARPS);
      mbar_init(&ch_full[b], FU_EPI_WARPS);
      mbar_init(&ch_empty[b], 1);
    }
    for (int s = 0; s < FU_MAX_RA; ++s) { mbar_init(&ra_full[s], 1); mbar_init(&ra_empty[s], 1); }
    for (int s = 0; s < FU_MAX_RB; ++s) { mbar_init(&rb_full[s], 1); mbar_init(&rb_empty[s], 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = threadIdx.x; i < p.N1; i += FU_THREADS) sba[i] = p.bias_a ? p.bias_a[i] : 0.f;
  for (int i = threadIdx.x; i < p.N2; i += FU_THREADS) sbb[i] = p.bias_b ? p.bias_b[i] : 0.f;
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int nq_rows = p.N2 / NQ;            // rows of W_b per RB box = columns of the output per B-MMA

  if (warp == 0) {
    // ---------------- producer 0: the tile's A0 operand, then the W_a boxes in consumption order
    const bool issuer = elect_one();
    int rs = 0;
    uint32_t rph = 0;
    int it = 0;
    for (int t = blockIdx.x; t < p.num_m_tiles; t += gridDim.x, ++it) {
      mbar_wait(a0_empty, (uint32_t)(it & 1) ^ 1);
      if (issuer) {
        mbar_expect_tx(a0_full, (uint32_t)KB0 * 16384u);
        for (int kb = 0; kb < KB0; ++kb) tma_load_2d(sA0 + kb * 16384, &tmA, a0_full, kb * 64, t * TC_BM);
      }
      __syncwarp();
      for (int j = 0; j < NC; ++j) {
        for (int kb = 0; kb < KB0; ++kb) {
          mbar_wait(&ra_empty[rs], rph ^ 1);
          if (issuer) {
            mbar_expect_tx(&ra_full[rs], FU_RA_BOX);
            tma_load_2d(sRA + rs * FU_RA_BOX, &tmWa, &ra_full[rs], kb * 64, j * FU_CHUNK);
          }
          __syncwarp();
          if (++rs == p.ra_slots) { rs = 0; rph ^= 1; }
        }
      }
    }
  } else if (warp == 2 + FU_EPI_WARPS) {
    // ---------------- producer 1: the W_b boxes (quarter of the output columns x one 64-wide K block)
    const bool issuer = elect_one();
    int rs = 0;
    uint32_t rph = 0;
    for (int t = blockIdx.x; t < p.num_m_tiles; t += gridDim.x) {
      for (int j = 0; j < NC; ++j) {
        for (int qd = 0; qd < NQ; ++qd) {
          mbar_wait(&rb_empty[rs], rph ^ 1);
          if (issuer) {
            mbar_expect_tx(&rb_full[rs], (uint32_t)p.rb_box);
            tma_load_2d(sRB + rs * p.rb_box, &tmWb, &rb_full[rs], j * FU_CHUNK, qd * nq_rows);
          }
          __syncwarp();
          if (++rs == p.rb_slots) { rs = 0; rph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ---------------- MMA issuer
    const bool issuer = elect_one();
    const uint64_t dconst = umma_desc_sw128(0);
    const uint32_t idesc_a = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(FU_CHUNK >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
    const uint32_t idesc_b = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(nq_rows >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
    const uint32_t a0_base = smem_u32(sA0) >> 4, ch_base = smem_u32(sCH) >> 4;
    const uint32_t ra_base = smem_u32(sRA) >> 4, rb_base = smem_u32(sRB) >> 4;
    int ras = 0, rbs = 0;
    uint32_t raph = 0, rbph = 0;
    uint32_t gc = 0;          // global chunk counter of the A-GEMM (selects the chunk accumulator and its parity)
    uint32_t gcb = 0;         // global chunk counter of the B-GEMM
    int it = 0;
    // A(j): chunk accumulator (gc & 1) = A0 . W_a[chunk j]^T
    auto issue_A = [&](bool last_of_tile) {
      const uint32_t b = gc & 1;
      mbar_wait(&acc2_empty[b], ((gc >> 1) & 1) ^ 1);
      tc_fence_after();
      const uint32_t d = tmem_base + FU_ACC2_COL + b * FU_CHUNK;
      for (int kb = 0; kb < KB0; ++kb) {
        mbar_wait(&ra_full[ras], raph);
        tc_fence_after();
        if (issuer) {
          const uint64_t ad = dconst | (uint64_t)(a0_base + kb * (16384 >> 4));
          const uint64_t bd = dconst | (uint64_t)(ra_base + ras * (FU_RA_BOX >> 4));
#pragma unroll
          for (int k = 0; k < 4; ++k) tc_mma(d, ad + 2 * k, bd + 2 * k, idesc_a, (uint32_t)((kb | k) != 0));
          tc_commit(&ra_empty[ras]);
        }
        __syncwarp();
        if (++ras == p.ra_slots) { ras = 0; raph ^= 1; }
      }
      if (issuer) {
        tc_commit(&acc2_full[b]);
        if (last_of_tile) tc_commit(a0_empty);   // every MMA that reads this tile's A0 has been issued before this commit
      }
      __syncwarp();
      ++gc;
    };
    for (int t = blockIdx.x; t < p.num_m_tiles; t += gridDim.x, ++it) {
      if (issuer) fu_trace(p, it, 0, 4, clock64());             // tile start
      mbar_wait(a0_full, (uint32_t)(it & 1));
      if (issuer) fu_trace(p, it, 0, 5, clock64());             // A0 present
      tc_fence_after();
      issue_A(NC == 1);
      if (issuer) fu_trace(p, it, 0, 6, clock64());             // A(0) issued
      for (int j = 0; j < NC; ++j) {
        if (j + 1 < NC) issue_A(j + 1 == NC - 1);
        // B(j): output accumulator += chunk(j) . W_b[:, chunk j]^T, one MMA group per quarter of the output columns
        const uint32_t b = gcb & 1;
        if (issuer) fu_trace(p, it, j, 0, clock64());           // A(j+1) issued
        mbar_wait(&ch_full[b], (gcb >> 1) & 1);
        if (issuer) fu_trace(p, it, j, 1, clock64());           // chunk j operand ready
        if (j == 0) mbar_wait(acc3_empty, (uint32_t)(it & 1) ^ 1);
        tc_fence_after();
        for (int qd = 0; qd < NQ; ++qd) {
          mbar_wait(&rb_full[rbs], rbph);
          if (issuer && qd == 0) fu_trace(p, it, j, 2, clock64());   // first W_b box of the chunk present
          if (issuer && qd == 3) fu_trace(p, it, j, 3, clock64());   // last W_b box present
          tc_fence_after();
          if (issuer) {
            const uint64_t ad = dconst | (uint64_t)(ch_base + b * (16384 >> 4));
            const uint64_t bd = dconst | (uint64_t)(rb_base + rbs * (p.rb_box >> 4));
            const uint32_t d = tmem_base + (uint32_t)(qd * nq_rows);
#pragma unroll
            for (int k = 0; k < 4; ++k) tc_mma(d, ad + 2 * k, bd + 2 * k, idesc_b, (uint32_t)((j | k) != 0));
            tc_commit(&rb_empty[rbs]);
          }
          __syncwarp();
          if (++rbs == p.rb_slots) { rbs = 0; rbph ^= 1; }
        }
        if (issuer) {
          fu_trace(p, it, j, 7, clock64());                       // B(j) issued
          tc_commit(&ch_empty[b]);
          if (j == NC - 1) tc_commit(acc3_full);
        }
        __syncwarp();
        ++gcb;
      }
    }
  } else {
    // ---------------- epilogue warps 2..9
    const int ew = warp - 2;
    const int q = warp & 3, sub = ew >> 2;      // row quarter, 16-column slice of a chunk
    const int h = sub;                          // tile epilogue: warps with sub < 2 take alternating 64-column groups
    const uint32_t lane_field = (uint32_t)(q * 32) << 16;
    const int trow = q * 32 + lane;                      // row inside the tile
    uint8_t* stg = sCH + (h & 1) * 16384 + q * 4096;     // store staging box of the warps with sub < 2
    uint32_t gc = 0;
    int it = 0;
    for (int t = blockIdx.x; t < p.num_m_tiles; t += gridDim.x, ++it) {
      const int row0 = t * TC_BM + q * 32;
      const int row = row0 + lane;
      const bool row_ok = row < p.M;
      const float* gb_row = (p.gbias && row0 < p.M) ? p.gbias + (size_t)(row0 / p.rows_per_group) * p.N1 : nullptr;
      // ---- chunk epilogues: chunk accumulator -> bias (+group bias) -> ReLU -> bf16 K-major smem operand.
      // The wait -> TMEM load -> convert -> store -> signal chain of a chunk sits on the critical path between the
      // two GEMMs, and one warp runs ~5 cycles per instruction of it; so every chunk is cut into 16-column slices
      // over 4 warps per row quarter (16 warps), ~80 instructions each.
      for (int j = 0; j < NC; ++j, ++gc) {
        const uint32_t b = gc & 1;
        const int c0 = j * FU_CHUNK + sub * 16;
        float gpre = 0.f;
        if (gb_row && lane < 16) gpre = __ldg(gb_row + c0 + lane);
        mbar_wait(&acc2_full[b], (gc >> 1) & 1);
        tc_fence_after();
        float v[16];
        tc_ld16_issue(tmem_base + lane_field + FU_ACC2_COL + b * FU_CHUNK + sub * 16, v);
        tc_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&acc2_empty[b]);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float4 b4 = *reinterpret_cast<const float4*>(sba + c0 + 4 * i);
          v[4 * i] += b4.x; v[4 * i + 1] += b4.y; v[4 * i + 2] += b4.z; v[4 * i + 3] += b4.w;
        }
        if (gb_row) {
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] += __shfl_sync(0xffffffffu, gpre, i);
        }
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = fmaxf(v[i], 0.f);
        mbar_wait(&ch_empty[b], ((gc >> 1) & 1) ^ 1);     // the B-GEMM that read this buffer two chunks ago is done
        const uint32_t rbase = smem_u32(sCH) + b * 16384 + trow * 128;
#pragma unroll
        for (int pc = 0; pc < 2; ++pc) {
          const uint32_t a = rbase + (((uint32_t)(pc + 2 * sub) ^ (trow & 7)) << 4);
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(pack_bf16x2(v[pc * 8], v[pc * 8 + 1])),
                       "r"(pack_bf16x2(v[pc * 8 + 2], v[pc * 8 + 3])), "r"(pack_bf16x2(v[pc * 8 + 4], v[pc * 8 + 5])),
                       "r"(pack_bf16x2(v[pc * 8 + 6], v[pc * 8 + 7]))
                       : "memory");
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the MMA (async proxy)
        __syncwarp();
        if (lane == 0) mbar_arrive(&ch_full[b]);
      }
      // ---- tile epilogue: output accumulator -> bias -> bf16 store and/or 32-row max
      mbar_wait(acc3_full, (uint32_t)(it & 1));
      tc_fence_after();
      const int ngroups = p.N2 / 64;
      for (int gi = h; gi < ngroups && h < 2; gi += 2) {
        const int n0 = gi * 64;
        const bool last = gi + 2 >= ngroups;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          float v[32];
          tc_ld32_issue(tmem_base + lane_field + (uint32_t)(n0 + half * 32), v);
          tc_ld_wait();
          if (last && half == 1) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(acc3_empty);      // the next tile's B-GEMM may overwrite the accumulator
          }
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float4 b4 = *reinterpret_cast<const float4*>(sbb + n0 + half * 32 + 4 * i);
            v[4 * i] += b4.x; v[4 * i + 1] += b4.y; v[4 * i + 2] += b4.z; v[4 * i + 3] += b4.w;
          }
          if (p.store_out) {
            const uint32_t rbase = smem_u32(stg) + lane * 128;
#pragma unroll
            for (int pc = 0; pc < 4; ++pc) {
              const uint32_t a = rbase + (((uint32_t)(pc + 4 * half) ^ (lane & 7)) << 4);
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(pack_bf16x2(v[pc * 8], v[pc * 8 + 1])),
                           "r"(pack_bf16x2(v[pc * 8 + 2], v[pc * 8 + 3])), "r"(pack_bf16x2(v[pc * 8 + 4], v[pc * 8 + 5])),
                           "r"(pack_bf16x2(v[pc * 8 + 6], v[pc * 8 + 7]))
                           : "memory");
            }
          }
          if (p.out_max || p.out_max_bf16) {
            if (!row_ok) {
#pragma unroll
              for (int i = 0; i < 32; ++i) v[i] = -3.0e38f;
            }
            float m = warp_rows_max(v, lane);
            if (p.max_relu) m = fmaxf(m, 0.f);
            const size_t o = (size_t)(row0 >> 5) * p.N2 + n0 + half * 32 + lane;
            if (row0 < p.M) {
              if (p.out_max) p.out_max[o] = m;
              if (p.out_max_bf16) p.out_max_bf16[o] = __float2bfloat16_rn(m);
            }
          }
        }
        if (p.store_out) {
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          __syncwarp();
          if (lane == 0 && row0 < p.M) {
            asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                             reinterpret_cast<uint64_t>(&tmC)),
                         "r"(smem_u32(stg)), "r"(n0), "r"(row0)
                         : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // single box per warp: reuse needs the read done
          }
          __syncwarp();
        }
      }
      if (h >= ngroups || h >= 2) {   // warps without an output group still hand the accumulator back
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(acc3_empty);
      }
      if (p.store_out) {
        // the staging boxes alias the chunk buffers: every warp's store must have left shared memory before any
        // warp starts the next tile's chunk epilogues
        asm volatile("bar.sync 1, %0;" ::"n"(FU_EPI_WARPS * 32) : "memory");
      }
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
  }
}

// ------------------------------------------------------------------------------------------------ host
static inline int make_map_box(CUtensorMap* m, const void* base, int64_t rows, int64_t cols, int box_rows) {
  return make_map(m, base, rows, cols, box_rows);
}

bool tc_fused_supported(int K0, int N1, int N2, int64_t rows_per_group, bool has_gbias) {
  if (K0 % 64 || N1 % 64 || N2 % 64) return false;
  if (K0 < 64 || K0 > 384 || N2 < 64 || N2 > 384 || N1 < 64 || N1 > 2048) return false;
  if (has_gbias && (rows_per_group % 32 != 0)) return false;
  return true;
}

// out = W_b relu(W_a A0 + bias_a + gbias) + bias_b.  A0 [M,K0] bf16, W_a [N1,K0], W_b [N2,N1] bf16.
int tc_fused(const __nv_bfloat16* A0, int64_t M, int K0, const __nv_bfloat16* Wa, int N1, const float* bias_a,
             const float* gbias, int rows_per_group, const __nv_bfloat16* Wb, int N2, const float* bias_b,
             __nv_bfloat16* out_bf16, float* out_max, __nv_bfloat16* out_max_bf16, int max_relu, cudaStream_t s) {
  P3_REQUIRE(tc_fused_supported(K0, N1, N2, rows_per_group, gbias != nullptr), P3TOK_ERR_UNSUPPORTED,
             "tc_fused: unsupported shape K0=%d N1=%d N2=%d", K0, N1, N2);
  P3_REQUIRE(M < (1ll << 31) - 256, P3TOK_ERR_UNSUPPORTED, "tc_fused: too many rows");
  if (M == 0) return P3TOK_OK;
  FusedParams p;
  p.M = (int)M; p.K0 = K0; p.N1 = N1; p.N2 = N2;
  p.num_m_tiles = (int)((M + TC_BM - 1) / TC_BM);
  p.bias_a = bias_a; p.gbias = gbias; p.rows_per_group = rows_per_group > 0 ? rows_per_group : 32; p.bias_b = bias_b;
  p.store_out = out_bf16 != nullptr; p.out_max = out_max; p.out_max_bf16 = out_max_bf16; p.max_relu = max_relu;
  p.rb_box = (N2 / 4) * 128;
  // shared-memory budget: what is left after the resident operands is split between the two weight rings
  const int fixed = (K0 / 64) * 16384 + 2 * 16384 + (N1 + N2 + 8 * 64) * 4 + 64 * 8 + 64 + 1024;
  int left = FU_SMEM - fixed;
  P3_REQUIRE(left >= 2 * FU_RA_BOX + 2 * p.rb_box, P3TOK_ERR_UNSUPPORTED, "tc_fused: shapes do not fit shared memory");
  // one chunk of W_a is K0/64 boxes of 8 KB, one chunk of W_b is 4 boxes of rb_box: give both the same number of chunks
  const int chunk_a = (K0 / 64) * FU_RA_BOX, chunk_b = 4 * p.rb_box;
  int ra_bytes = (int)((int64_t)left * chunk_a / (chunk_a + chunk_b));
  p.ra_slots = ra_bytes / FU_RA_BOX;
  if (p.ra_slots > FU_MAX_RA) p.ra_slots = FU_MAX_RA;
  if (p.ra_slots < 2) p.ra_slots = 2;
  p.rb_slots = (left - p.ra_slots * FU_RA_BOX) / p.rb_box;
  if (p.rb_slots > FU_MAX_RB) p.rb_slots = FU_MAX_RB;
  P3_REQUIRE(p.rb_slots >= 2, P3TOK_ERR_UNSUPPORTED, "tc_fused: shapes do not fit shared memory");
  CUtensorMap ta, twa, twb, tc;
  int rc = make_map_box(&ta, A0, M, K0, TC_BM);
  if (rc) return rc;
  rc = make_map_box(&twa, Wa, N1, K0, FU_CHUNK);
  if (rc) return rc;
  rc = make_map_box(&twb, Wb, N2, N1, N2 / 4);
  if (rc) return rc;
  if (out_bf16) {
    rc = make_map_box(&tc, out_bf16, M, N2, 32);
    if (rc) return rc;
  } else {
    tc = ta;
  }
  static thread_local bool configured[32] = {false};
  int dev = 0;
  P3_CUDA(cudaGetDevice(&dev));
  if (dev < 32 && !configured[dev]) {
    P3_CUDA(cudaFuncSetAttribute(tc_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FU_SMEM));
    configured[dev] = true;
  }
  static int trace_on = -1;
  if (trace_on < 0) trace_on = getenv("P3TOK_TC_TRACE") ? 1 : 0;
  p.trace = nullptr;
  const size_t tw = 4 * 16 * 8;
  if (trace_on) {
    P3_CUDA(cudaMalloc(&p.trace, tw * 8));
    P3_CUDA(cudaMemsetAsync(p.trace, 0, tw * 8, s));
  }
  const int grid = p.num_m_tiles < num_sms() ? p.num_m_tiles : num_sms();
  tc_fused_kernel<<<grid, FU_THREADS, FU_SMEM, s>>>(ta, twa, twb, tc, p);
  P3_LAUNCH_CHECK("tc_fused_kernel");
  if (trace_on) {   // debug only: synchronises
    std::vector<unsigned long long> h(tw);
    P3_CUDA(cudaStreamSynchronize(s));
    P3_CUDA(cudaMemcpy(h.data(), p.trace, tw * 8, cudaMemcpyDeviceToHost));
    P3_CUDA(cudaFree(p.trace));
    fprintf(stderr, "[fu_trace] M=%d K0=%d N1=%d N2=%d ra=%d rb=%d\n", p.M, p.K0, p.N1, p.N2, p.ra_slots, p.rb_slots);
    const unsigned long long t0 = h[4];
    for (int it = 0; it < 3; ++it) {
      const unsigned long long* q = &h[(size_t)it * 16 * 8];
      auto rel = [&](unsigned long long v) { return v ? (long long)(v - t0) : -1ll; };
      fprintf(stderr, "[fu_trace] tile%d start=%lld a0_ok=%lld A0_issued=%lld\n", it, rel(q[4]), rel(q[5]), rel(q[6]));
      for (int j = 0; j < 16 && q[j * 8 + 7]; ++j)
        fprintf(stderr, "[fu_trace]   chunk%-2d A_next_issued=%lld ch_ok=%lld wb_first=%lld wb_last=%lld B_issued=%lld\n", j,
                rel(q[j * 8 + 0]), rel(q[j * 8 + 1]), rel(q[j * 8 + 2]), rel(q[j * 8 + 3]), rel(q[j * 8 + 7]));
    }
  }
  return P3TOK_OK;
}

}  // namespace p3tok

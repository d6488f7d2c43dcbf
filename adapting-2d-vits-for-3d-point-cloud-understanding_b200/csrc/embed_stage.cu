// embed_stage.cu - the concat layer and the output layer of a NARROW patch-embedding block in one tcgen05
// kernel whose weights never leave shared memory (CTA pairs, cta_group::2).
//
//   token[g] = max over the 32 rows of patch g of  act( W_b . relu(W_a . a0[row] + gbias[g]) + bias_b )
//
// This is conv2 + pool of a P3Embed stage (reference src/models/pix4point.py:147-156, 184-188) when the stage is
// narrow enough that both folded matrices fit next to the activations: stage 0 of the 2-stage embedding BASELINE
// configs 1/3/5 use (W = 128: a0 128 wide, hidden 256, output 128) and the W = 64 first stage of a 3-stage one.
// gbias[g] = W_g . max_k(a0) + b is the pooled half of the concat layer (pix4point.py:184-186), produced by a
// tensor-core GEMM over groups (tc_linear) beforehand.
//
// Why a separate kernel: at these widths the layer-by-layer path is HBM-bound by a factor of 4 (per 2^20 rows it
// moves 0.27 + 0.54 GB out and 0.27 + 0.54 GB back in for 0.14 TFLOP: 563 us measured vs 98 us of tensor time),
// and the generic two-layer kernel (embed_fused.cu) streams weight boxes and is limited to 64-column chunks
// (~90 cycles per tcgen05.mma however small N is).  Here the hidden activation lives in shared memory, the weights
// are loaded ONCE per CTA, and every MMA is full width (N = hidden width / output width, M = 256 across the pair).
//
// Per CTA (its 128 rows of the pair's 256-row tile), shared memory:
//   W_a half  [N1/2 rows][K0]   K-major, 128B swizzle, one 64-column block after the other     (32 KB)
//   W_b half  [N2/2 rows][N1]                                                                  (32 KB)
//   a0 tiles  2 stages x [128][K0]      TMA ring, producer one tile ahead                      (64 KB)
//   hidden    [128][N1] bf16            written by the hidden epilogue, read by the 2nd GEMM   (64 KB)
// TMEM (512 columns): two accumulators of 256 columns.  Tile t uses accumulator t&1: GEMM 1 fills columns [0,N1);
// once the hidden epilogue has drained them, GEMM 2 of the same tile re-uses columns [0,N2) for the output.
// Warps: 0 = a0 producer (TMA), 1 = MMA issuer (leader CTA only), 2-9 = hidden epilogue (row quarter q x column
// half h), 10-17 = output epilogue (row quarter q x column half: 32-row max, bias, ReLU, token store).
// MMA issue order  G1(0) G1(1) | G2(0) G1(2) | G2(1) G1(3) ...: while the hidden epilogue of tile t+1 runs, the
// tensor pipe executes G2(t) and G1(t+2).
// Barriers the leader's MMA warp waits on live in the leader CTA (TMA loads of both CTAs signal them, epilogue warps
// of the peer arrive remotely); barriers producers / epilogue warps wait on are local and are released by
// multicast tcgen05.commit.
// Measured and dropped (round 2): forming a0 INSIDE this kernel - the output-epilogue warps computing the 6 -> 128 first
// layer from 32 bytes of gathered inputs per row straight into the swizzled a0 stage, like the APF first layer in
// embed_fused.cu - so that rows_first_layer_narrow_kernel only has to emit the patch max.  Bit-identical tokens, but slower:
// c5w 2.02 -> 2.24 ms, c3 10.65 -> 11.65 ms per step.  A tile of this kernel is only ~3650 cycles and already bound by shared-
// memory bandwidth and by the hidden epilogue's issue slots; ~4300 more warp instructions per tile (30 % of the issue slots
// of its 3650 cycles) cost more than the 148 us / 2^21 rows of the separate first-layer kernel they replace.  (The pair
// kernel absorbs its first layer because ITS tile is 10-25 k cycles.)
#include "tc_common.cuh"

namespace p3tok {

#ifndef P3TOK_ST_E1_WARPS
#define P3TOK_ST_E1_WARPS 8
#endif
constexpr int ST_E1_WARPS = P3TOK_ST_E1_WARPS, ST_E2_WARPS = 8;
constexpr int ST_THREADS = (2 + ST_E1_WARPS + ST_E2_WARPS) * 32;
constexpr int ST_SMEM = 227 * 1024;
constexpr int ST_ACC_COLS = 256;            // TMEM columns per accumulator

struct StageParams {
  int M, K0, N1, N2;
  int num_pairs;                    // ceil(num_m_tiles / 2): one 256-row tile per CTA pair
  const float* gbias;               // [M / rows_per_group, N1] (bias of the hidden layer included) or null
  const float* bias_a;              // [N1] used when gbias is null, or null
  int rows_per_group;               // multiple of 32
  const float* bias_b;              // [N2] or null
  float* out_max;                   // [ceil(M/32), N2] f32 and / or
  __nv_bfloat16* out_max_bf16;      // the same rounded to bf16 (either may be null)
  int max_relu;
  unsigned long long* trace;        // debug (P3TOK_TC_TRACE=1): leader CTA of pair 0, 16 clock stamps per tile
};
// slots: MMA warp 0 g2-wait 1 h_full-ok 2 g2-issued 3 g1-wait 4 a0_full-ok 5 acc_free-ok 6 g1-issued | hidden epilogue (warp 2)
// 7 wait 8 acc1_full-ok 9 h_empty-ok 10 published | output epilogue (warp 10) 11 wait 12 acc2_full-ok 13 released 14 done | 15 TMA issued
__device__ __forceinline__ void st_trace(const StageParams& p, int it, int slot) {
  if (p.trace && blockIdx.x == 0 && it < 32) p.trace[it * 16 + slot] = (unsigned long long)clock64();
}

// One 32-row x 32-column piece of the hidden tile: accumulator + group bias -> ReLU -> bf16x2 (16 registers) ...
__device__ __forceinline__ void st_convert(const float (&v)[32], const float* gb32, uint32_t (&pk)[16]) {
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float4 g4 = *reinterpret_cast<const float4*>(gb32 + 4 * i);
    float a0, a1, a2, a3;
    add2(a0, a1, v[4 * i], v[4 * i + 1], g4.x, g4.y);
    add2(a2, a3, v[4 * i + 2], v[4 * i + 3], g4.z, g4.w);
    pk[2 * i] = pack_bf16x2_relu(a0, a1);          // ReLU rides on the conversion
    pk[2 * i + 1] = pack_bf16x2_relu(a2, a3);
  }
}
// ... and its store into the swizzled K-major rows of the hidden tile (64 B of this thread's 128-byte row)
__device__ __forceinline__ void st_store(const uint32_t (&pk)[16], uint32_t rbase, int h, int trow) {
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const uint32_t a = rbase + (((uint32_t)(j + 4 * h) ^ (uint32_t)(trow & 7)) << 4);
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(pk[4 * j]), "r"(pk[4 * j + 1]), "r"(pk[4 * j + 2]),
                 "r"(pk[4 * j + 3])
                 : "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the MMA (async proxy)
}

__global__ void __launch_bounds__(ST_THREADS, 1)
tc_stage_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmWa,
                const __grid_constant__ CUtensorMap tmWb, const StageParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int KB0 = p.K0 / 64, KB1 = p.N1 / 64;
  const int wa_block = (p.N1 / 2) * 128, wb_block = (p.N2 / 2) * 128;   // bytes per 64-column K block of this CTA's half
  const int a0_bytes = KB0 * 16384;
  uint8_t* sWa = smem;
  uint8_t* sWb = sWa + KB0 * wa_block;
  uint8_t* sA0 = sWb + KB1 * wb_block;                 // 2 stages
  uint8_t* sH = sA0 + 2 * a0_bytes;                    // KB1 x 16 KB
  float* sgb = reinterpret_cast<float*>(sH + KB1 * 16384);   // E1 warps x 128 floats (group-bias slices)
  float* sbb = sgb + ST_E1_WARPS * 128;                // N2 floats
  uint64_t* bars = reinterpret_cast<uint64_t*>(sbb + 256);
  uint64_t* w_full = bars;             // leader
  uint64_t* a0_full = bars + 1;        // [2] leader
  uint64_t* a0_empty = bars + 3;       // [2] local, multicast commit
  uint64_t* acc1_full = bars + 5;      // [2] local, multicast commit
  uint64_t* acc2_full = bars + 7;      // [2] local, multicast commit
  uint64_t* acc_free = bars + 9;       // [2] leader, 2 x ST_E2_WARPS arrivals
  uint64_t* h_full = bars + 11;        // leader, 2 x ST_E1_WARPS arrivals
  uint64_t* h_empty = bars + 12;       // local, multicast commit
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 13);

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const int rank = (int)cluster_ctarank();
  const int pair_id = blockIdx.x >> 1, pair_stride = gridDim.x >> 1;

  if (threadIdx.x == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmA)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmWa)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmWb)) : "memory");
    mbar_init(w_full, 1);
    for (int b = 0; b < 2; ++b) {
      mbar_init(&a0_full[b], 1);
      mbar_init(&a0_empty[b], 1);
      mbar_init(&acc1_full[b], 1);
      mbar_init(&acc2_full[b], 1);
      mbar_init(&acc_free[b], 2 * ST_E2_WARPS);
    }
    mbar_init(h_full, 2 * ST_E1_WARPS);
    mbar_init(h_empty, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = threadIdx.x; i < p.N2; i += ST_THREADS) sbb[i] = p.bias_b ? p.bias_b[i] : 0.f;
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::);
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                       // the peer's barriers exist before any remote arrive / multicast commit
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ---------------- producer: this CTA's half of both weight matrices (once), then its a0 tiles
    const bool issuer = elect_one();
    if (issuer) {
      if (rank == 0) mbar_expect_tx(w_full, 2u * (uint32_t)(KB0 * wa_block + KB1 * wb_block));
      for (int kb = 0; kb < KB0; ++kb) tma_load_2d_pair(sWa + kb * wa_block, &tmWa, w_full, kb * 64, rank * (p.N1 / 2));
      for (int kb = 0; kb < KB1; ++kb) tma_load_2d_pair(sWb + kb * wb_block, &tmWb, w_full, kb * 64, rank * (p.N2 / 2));
    }
    __syncwarp();
    int it = 0;
    for (int tp = pair_id; tp < p.num_pairs; tp += pair_stride, ++it) {
      const int s = it & 1;
      const int mt = 2 * tp + rank;         // may be a dummy tile past the end: TMA zero-fills, stores are clipped
      mbar_wait(&a0_empty[s], ((uint32_t)(it >> 1) & 1) ^ 1);
      if (issuer) {
        if (rank == 0) mbar_expect_tx(&a0_full[s], 2u * (uint32_t)a0_bytes);
        for (int kb = 0; kb < KB0; ++kb) tma_load_2d_pair(sA0 + s * a0_bytes + kb * 16384, &tmA, &a0_full[s], kb * 64, mt * TC_BM);
        st_trace(p, it, 15);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    if (rank == 0) {
      // ---------------- MMA issuer (leader CTA, for both SMs of the pair)
      const bool issuer = elect_one();
      const uint64_t dconst = umma_desc_sw128(0);
      const uint32_t idesc1 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.N1 >> 3) << 17) | ((uint32_t)((2 * TC_BM) >> 4) << 24);
      const uint32_t idesc2 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.N2 >> 3) << 17) | ((uint32_t)((2 * TC_BM) >> 4) << 24);
      const uint32_t a0_base = smem_u32(sA0) >> 4, h_base = smem_u32(sH) >> 4;
      const uint32_t wa_base = smem_u32(sWa) >> 4, wb_base = smem_u32(sWb) >> 4;
      int n_tiles = 0;
      for (int tp = pair_id; tp < p.num_pairs; tp += pair_stride) ++n_tiles;
      // G1(t): accumulator t&1, columns [0,N1) = a0(t) . W_a^T
      auto issue_g1 = [&](int t) {
        const int s = t & 1;
        if (issuer) st_trace(p, t, 3);
        mbar_wait(&a0_full[s], (uint32_t)(t >> 1) & 1);
        if (issuer) st_trace(p, t, 4);
        mbar_wait(&acc_free[s], ((uint32_t)(t >> 1) & 1) ^ 1);     // output epilogue of tile t-2 has drained this accumulator
        tc_fence_after();
        if (issuer) {
          st_trace(p, t, 5);
          const uint32_t d = tmem_base + (uint32_t)(s * ST_ACC_COLS);
          for (int kb = 0; kb < KB0; ++kb) {
            const uint64_t ad = dconst | (uint64_t)(a0_base + ((s * a0_bytes + kb * 16384) >> 4));
            const uint64_t bd = dconst | (uint64_t)(wa_base + ((kb * wa_block) >> 4));
#pragma unroll
            for (int k = 0; k < 4; ++k) tc_mma_pair(d, ad + 2 * k, bd + 2 * k, idesc1, (uint32_t)((kb | k) != 0));
          }
          tc_commit_pair(&a0_empty[s]);
          tc_commit_pair(&acc1_full[s]);
          st_trace(p, t, 6);
        }
        __syncwarp();
      };
      // G2(t): accumulator t&1, columns [0,N2) = hidden(t) . W_b^T   (hidden epilogue has drained GEMM 1's result)
      auto issue_g2 = [&](int t) {
        const int s = t & 1;
        if (issuer) st_trace(p, t, 0);
        mbar_wait(h_full, (uint32_t)t & 1);
        tc_fence_after();
        if (issuer) {
          st_trace(p, t, 1);
          const uint32_t d = tmem_base + (uint32_t)(s * ST_ACC_COLS);
          for (int kb = 0; kb < KB1; ++kb) {
            const uint64_t ad = dconst | (uint64_t)(h_base + ((kb * 16384) >> 4));
            const uint64_t bd = dconst | (uint64_t)(wb_base + ((kb * wb_block) >> 4));
#pragma unroll
            for (int k = 0; k < 4; ++k) tc_mma_pair(d, ad + 2 * k, bd + 2 * k, idesc2, (uint32_t)((kb | k) != 0));
          }
          tc_commit_pair(h_empty);
          tc_commit_pair(&acc2_full[s]);
          st_trace(p, t, 2);
        }
        __syncwarp();
      };
      mbar_wait(w_full, 0);
      tc_fence_after();
      if (n_tiles > 0) issue_g1(0);
      if (n_tiles > 1) issue_g1(1);
      for (int t = 0; t < n_tiles; ++t) {
        issue_g2(t);
        if (t + 2 < n_tiles) issue_g1(t + 2);
      }
    }
  } else if (warp < 2 + ST_E1_WARPS) {
    // ---------------- hidden epilogue: GEMM 1 accumulator -> + group bias -> ReLU -> bf16 K-major operand of GEMM 2
    // (row quarter q x column half h).  Finer hand-over (one barrier per 64-column K block, TMEM loads one block ahead,
    // one warp per quarter) was built and measured no better - see the header; 16 hidden-epilogue warps
    // (-DP3TOK_ST_E1_WARPS=16) measured 6 % slower: the kernel is bound by shared-memory bandwidth, not by warp latency.
    const int ew = warp - 2;
    const int q = warp & 3, h = ew >> 2;            // TMEM lane quarter, column half
    const uint32_t lane_field = (uint32_t)(q * 32) << 16;
    const int trow = q * 32 + lane;                 // row inside this CTA's 128-row tile
    const int half_cols = p.N1 / (ST_E1_WARPS / 4), pieces = half_cols / 32;   // columns per warp of a row quarter
    float* my_sgb = sgb + ew * 128;
    const uint32_t h_row = smem_u32(sH) + trow * 128;
    // this warp's slice of the group-bias row of tile `tp`, one float4 per lane, fetched ONE TILE AHEAD (otherwise an L2
    // round trip per tile sits on the critical path between the two GEMMs)
    auto load_gb = [&](int tp) {
      float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
      if (tp >= p.num_pairs) return g;
      const int row0 = (2 * tp + rank) * TC_BM + q * 32;
      const float* src = nullptr;
      if (p.gbias) { if (row0 < p.M) src = p.gbias + (size_t)(row0 / p.rows_per_group) * p.N1 + h * half_cols; }
      else if (p.bias_a) src = p.bias_a + h * half_cols;
      if (src && 4 * lane < half_cols) g = __ldg(reinterpret_cast<const float4*>(src) + lane);
      return g;
    };
    float4 gnext = load_gb(pair_id);
    int it = 0;
    for (int tp = pair_id; tp < p.num_pairs; tp += pair_stride, ++it) {
      const int s = it & 1;
      const float4 gpre = gnext;
      const bool tr = (warp == 2 && lane == 0);
      if (tr) st_trace(p, it, 7);
      mbar_wait(&acc1_full[s], (uint32_t)(it >> 1) & 1);
      if (tr) st_trace(p, it, 8);
      tc_fence_after();
      __syncwarp();                                  // previous tile's reads of the slice are done
      *reinterpret_cast<float4*>(my_sgb + 4 * lane) = gpre;
      __syncwarp();
      gnext = load_gb(tp + pair_stride);
      for (int pc = 0; pc < pieces; ++pc) {
        const int c0 = h * half_cols + pc * 32;      // hidden column of v[0]
        float v[32];
        uint32_t pk[16];
        tc_ld32_issue(tmem_base + lane_field + (uint32_t)(s * ST_ACC_COLS + c0), v);
        tc_ld_wait();
        st_convert(v, my_sgb + pc * 32, pk);
        if (pc == 0) {
          mbar_wait(h_empty, ((uint32_t)it & 1) ^ 1);   // GEMM 2 of the previous tile has read the hidden tile
          if (tr) st_trace(p, it, 9);
        }
        st_store(pk, h_row + (uint32_t)(c0 >> 6) * 16384u, (c0 >> 5) & 1, trow);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cta(h_full, 0);
      if (tr) st_trace(p, it, 10);
    }
  } else {
    // ---------------- output epilogue: GEMM 2 accumulator -> max over the warp's 32 rows -> + bias -> ReLU -> token
    const int ew = warp - 2 - ST_E1_WARPS;
    const int q = warp & 3, h = ew >> 2;
    const uint32_t lane_field = (uint32_t)(q * 32) << 16;
    const int half_cols = p.N2 / (ST_E2_WARPS / 4), pieces = half_cols / 32;   // columns per warp of a row quarter
    int it = 0;
    for (int tp = pair_id; tp < p.num_pairs; tp += pair_stride, ++it) {
      const int s = it & 1;
      const int row0 = (2 * tp + rank) * TC_BM + q * 32;
      const bool tr = (warp == 2 + ST_E1_WARPS && lane == 0);
      if (tr) st_trace(p, it, 11);
      mbar_wait(&acc2_full[s], (uint32_t)(it >> 1) & 1);
      if (tr) st_trace(p, it, 12);
      tc_fence_after();
      // 32-row max straight from the m16n8 accumulator fragments (rows_max_frag, tc_common.cuh: 7 shuffles per 32 columns
      // instead of 32 redux.sync); the bias is added after the max (exactly the same value: fp32 rounding is monotonic)
      const int nvalid = p.M - row0 >= 32 ? 32 : (p.M - row0 > 0 ? p.M - row0 : 0);
      const int fcol = rows_max_frag_col(lane);
      for (int pc0 = 0; pc0 < pieces; pc0 += 2) {
        const bool two = pc0 + 1 < pieces;
        const int c0 = h * half_cols + pc0 * 32;
        float a0[16], b0[16], a1[16], b1[16];
        const uint32_t t0 = tmem_base + lane_field + (uint32_t)(s * ST_ACC_COLS + c0);
        tc_ld16x256_x4_issue(t0, a0);
        tc_ld16x256_x4_issue(t0 + (16u << 16), b0);
        if (two) {
          tc_ld16x256_x4_issue(t0 + 32u, a1);
          tc_ld16x256_x4_issue(t0 + 32u + (16u << 16), b1);
        }
        tc_ld_wait();
        if (pc0 + 2 >= pieces) {        // last accumulator read of this warp: GEMM 1 of tile t+2 may overwrite it
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_cta(&acc_free[s], 0);
          if (tr) st_trace(p, it, 13);
        }
        float m = rows_max_frag(a0, b0, lane, nvalid) + sbb[c0 + fcol];
        if (p.max_relu) m = fmaxf(m, 0.f);
        if (row0 < p.M) {
          const size_t o = (size_t)(row0 >> 5) * p.N2 + c0 + fcol;
          if (p.out_max) p.out_max[o] = m;
          if (p.out_max_bf16) p.out_max_bf16[o] = __float2bfloat16_rn(m);
        }
        if (two) {
          float m1 = rows_max_frag(a1, b1, lane, nvalid) + sbb[c0 + 32 + fcol];
          if (p.max_relu) m1 = fmaxf(m1, 0.f);
          if (row0 < p.M) {
            const size_t o = (size_t)(row0 >> 5) * p.N2 + c0 + 32 + fcol;
            if (p.out_max) p.out_max[o] = m1;
            if (p.out_max_bf16) p.out_max_bf16[o] = __float2bfloat16_rn(m1);
          }
        }
      }
      if (tr) st_trace(p, it, 14);
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();   // no CTA leaves while the peer may still signal it or read its operands
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
  }
}

// ------------------------------------------------------------------------------------------------ host
static int stage_smem_bytes(int K0, int N1, int N2) {
  return (K0 / 64) * (N1 / 2) * 128 + (N1 / 64) * (N2 / 2) * 128 + 2 * (K0 / 64) * 16384 + (N1 / 64) * 16384 +
         ST_E1_WARPS * 512 + 1024 + 16 * 8 + 64 + 1024;
}

bool tc_stage_supported(int K0, int N1, int N2, int64_t rows_per_group) {
  if (K0 % 64 || N1 % 64 || N2 % 64) return false;
  if (K0 < 64 || N1 < 64 || N2 < 64 || N1 > ST_ACC_COLS || N2 > N1) return false;
  if (rows_per_group % 32 != 0 || rows_per_group <= 0) return false;
  return stage_smem_bytes(K0, N1, N2) <= ST_SMEM;
}

// out_max[g] = max over each 32 rows of act(W_b relu(W_a A0 + gbias|bias_a) + bias_b).  A0 [M,K0] bf16, W_a [N1,K0], W_b [N2,N1].
int tc_stage(const __nv_bfloat16* A0, int64_t M, int K0, const __nv_bfloat16* Wa, int N1, const float* bias_a,
             const float* gbias, int rows_per_group, const __nv_bfloat16* Wb, int N2, const float* bias_b, float* out_max,
             __nv_bfloat16* out_max_bf16, int max_relu, cudaStream_t s) {
  if (rows_per_group <= 0) rows_per_group = 32;
  P3_REQUIRE(tc_stage_supported(K0, N1, N2, rows_per_group), P3TOK_ERR_UNSUPPORTED, "tc_stage: unsupported shape K0=%d N1=%d N2=%d",
             K0, N1, N2);
  P3_REQUIRE(M < (1ll << 31) - 512, P3TOK_ERR_UNSUPPORTED, "tc_stage: too many rows");
  P3_REQUIRE(out_max != nullptr || out_max_bf16 != nullptr, P3TOK_ERR_INVALID, "tc_stage: null output");
  if (M == 0) return P3TOK_OK;
  StageParams p;
  p.M = (int)M; p.K0 = K0; p.N1 = N1; p.N2 = N2;
  const int num_m_tiles = (int)((M + TC_BM - 1) / TC_BM);
  p.num_pairs = (num_m_tiles + 1) / 2;
  p.gbias = gbias; p.bias_a = bias_a; p.rows_per_group = rows_per_group; p.bias_b = bias_b;
  p.out_max = out_max; p.out_max_bf16 = out_max_bf16; p.max_relu = max_relu;
  CUtensorMap ta, twa, twb;
  int rc = make_map(&ta, A0, M, K0, TC_BM);
  if (rc) return rc;
  rc = make_map(&twa, Wa, N1, K0, N1 / 2);       // each CTA of the pair keeps half of the hidden rows of W_a
  if (rc) return rc;
  rc = make_map(&twb, Wb, N2, N1, N2 / 2);       // ... and half of the output rows of W_b
  if (rc) return rc;
  static thread_local bool configured[32] = {false};
  int dev = 0;
  P3_CUDA(cudaGetDevice(&dev));
  if (dev < 32 && !configured[dev]) {
    P3_CUDA(cudaFuncSetAttribute(tc_stage_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ST_SMEM));
    configured[dev] = true;
  }
  const int max_pairs = num_sms() / 2;
  const int pairs = p.num_pairs < max_pairs ? p.num_pairs : max_pairs;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(pairs * 2));
  cfg.blockDim = dim3(ST_THREADS);
  cfg.dynamicSmemBytes = (size_t)stage_smem_bytes(K0, N1, N2);
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  static int trace_on = -1;
  if (trace_on < 0) trace_on = getenv("P3TOK_TC_TRACE") ? 1 : 0;
  p.trace = nullptr;
  if (trace_on) {
    P3_CUDA(cudaMalloc(&p.trace, 32 * 16 * 8));
    P3_CUDA(cudaMemsetAsync(p.trace, 0, 32 * 16 * 8, s));
  }
  P3_CUDA(cudaLaunchKernelEx(&cfg, tc_stage_kernel, ta, twa, twb, p));
  count_launch();
  if (trace_on) {   // debug only: synchronises and prints the leader CTA of pair 0 (cycles relative to its first stamp)
    unsigned long long h[32 * 16];
    P3_CUDA(cudaStreamSynchronize(s));
    P3_CUDA(cudaMemcpy(h, p.trace, sizeof(h), cudaMemcpyDeviceToHost));
    P3_CUDA(cudaFree(p.trace));
    const unsigned long long t0 = h[3];
    auto rel = [&](unsigned long long v) { return v ? (long long)(v - t0) : -1ll; };
    fprintf(stderr, "[st_trace] M=%d K0=%d N1=%d N2=%d pairs=%d\n", p.M, K0, N1, N2, pairs);
    for (int it = 0; it < 32 && h[it * 16 + 3]; ++it) {
      const unsigned long long* q = &h[it * 16];
      fprintf(stderr, "[st_trace] t%-2d g1[wait=%lld a0=%lld accfree=%lld issued=%lld] g2[wait=%lld hfull=%lld issued=%lld] "
              "e1[wait=%lld acc1=%lld hempty=%lld pub=%lld] e2[wait=%lld acc2=%lld rel=%lld done=%lld] tma=%lld\n", it,
              rel(q[3]), rel(q[4]), rel(q[5]), rel(q[6]), rel(q[0]), rel(q[1]), rel(q[2]), rel(q[7]), rel(q[8]), rel(q[9]),
              rel(q[10]), rel(q[11]), rel(q[12]), rel(q[13]), rel(q[14]), rel(q[15]));
    }
  }
  return P3TOK_OK;
}

}  // namespace p3tok

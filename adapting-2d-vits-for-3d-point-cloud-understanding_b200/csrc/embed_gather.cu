// embed_gather.cu - the first layer of a WIDE-input patch-embedding block with the gather fused into the GEMM's producer
// (CTA pairs, cta_group::2).
//
//   h[row] = relu( W . [feats[b, nbr(row)] || xyz[b, nbr(row)]] + bias )        (bf16 out)   and   g[group] = max over its k rows
//
// This is conv1 (both activation-free convolutions folded into one matrix, p3tok/fold.py) + BN + ReLU of a P3Embed stage
// >= 1 (reference src/models/pix4point.py:179-186: `group_knn` gather, concat [dp, fj], conv1, max over k), input width
// 3 + D with D = the previous stage's width (128 at BASELINE configs 1/3/5).  Round 1 materialised the gathered rows
// (rows_gather_p4p_bf16_kernel: 204 us per 2^20 rows, 285 MB written and read back) and ran tc_linear on them (226 us,
// bound by its 537 MB output); here 8 producer warps gather feature rows from L2 straight into the swizzled K-major A tile
// - float4 loads (a half-warp per row and 64-column k-block), bf16 conversion, 8-byte shared stores at the swizzled
// position - while the tensor pipe works on the previous tile: the row matrix never exists.
//
// Column order of the A tile: [feats (D) | xyz (3) | zero pad to 64] (the host rotates / pads the weight columns to
// match, pad_weight_kernel rot = 3); the xyz k-block's 61 pad columns are zeroed once per kernel, only its first 16-byte
// chunk is rewritten per tile.
// Per CTA (its 128 rows of the pair's 256-row tile), shared memory:
//   W half   [N1/2 rows][KB x 64]  resident, loaded once          (<= 32 KB per k-block-pair ... KB * N1/2 * 128 B)
//   A tiles  2 buffers x KB x 16 KB, producers one tile ahead
//   store staging 8 warps x 4 KB
// TMEM: two accumulators of N1 (<= 256) columns.  Warps: 0 = weight loader, 1 = MMA issuer (leader CTA), 2-9 = epilogue
// (row quarter q x column half: bias, ReLU, bf16 box -> TMA store, patch max read back from the box), 10-17 = gather
// producers (16 rows each).  Barriers the leader's MMA warp waits on live in the leader CTA (the peer's producers arrive
// remotely after a proxy fence, exactly like the chunk operands of embed_fused.cu).
#include "tc_common.cuh"

namespace p3tok {

constexpr int GL_EPI_WARPS = 8, GL_PROD_WARPS = 8;
constexpr int GL_THREADS = (2 + GL_EPI_WARPS + GL_PROD_WARPS) * 32;
constexpr int GL_MAX_KB = 5;              // D <= 256 feature columns + the xyz block

struct GatherParams {
  int64_t M;                        // rows = groups * k of this chunk
  int64_t g_begin;                  // first group of the chunk (global numbering b * G + g)
  int N1, D, KB;                    // output width, feature width (multiple of 64), k-blocks = D / 64 + 1
  int num_pairs;
  int64_t N, G, k;                  // points per cloud, groups per cloud, rows per group
  const float* pts;                 // (B, N, 3)
  const float* feats;               // (B, N, D) channel-last
  const void* knn;                  // (B, G, k) int32 or int64
  int idx64;
  const float* bias;                // [N1] or null
  int relu;
  __nv_bfloat16* out_max_bf16;      // [M / 32, N1] max over every 32 rows, or null
};

template <typename IdxT, int DB>     // DB = feature k-blocks (D / 64): compile-time, the gathered rows live in registers
__global__ void __launch_bounds__(GL_THREADS, 1)
tc_gather_linear_kernel(const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmC, const GatherParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int KB = p.KB;
  const int w_block = (p.N1 / 2) * 128;                // bytes per 64-column K block of this CTA's half of W
  const int a_bytes = KB * 16384;
  uint8_t* sW = smem;                                  // KB x w_block
  uint8_t* sA = sW + KB * w_block;                     // 2 x a_bytes
  uint8_t* sST = sA + 2 * a_bytes;                     // 8 x 4 KB
  float* sb = reinterpret_cast<float*>(sST + GL_EPI_WARPS * 4096);   // N1 floats
  uint64_t* bars = reinterpret_cast<uint64_t*>(sb + 256);
  uint64_t* w_full = bars;             // leader
  uint64_t* a_full = bars + 1;         // [2] leader, 2 x GL_PROD_WARPS arrivals
  uint64_t* a_empty = bars + 3;        // [2] local, multicast commit
  uint64_t* acc_full = bars + 5;       // [2] local, multicast commit
  uint64_t* acc_free = bars + 7;       // [2] leader, 2 x GL_EPI_WARPS arrivals
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 9);

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const int rank = (int)cluster_ctarank();
  const int pair_id = blockIdx.x >> 1, pair_stride = gridDim.x >> 1;

  if (threadIdx.x == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmW)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmC)) : "memory");
    mbar_init(w_full, 1);
    for (int b = 0; b < 2; ++b) {
      mbar_init(&a_full[b], 2 * GL_PROD_WARPS);
      mbar_init(&a_empty[b], 1);
      mbar_init(&acc_full[b], 1);
      mbar_init(&acc_free[b], 2 * GL_EPI_WARPS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = threadIdx.x; i < p.N1; i += GL_THREADS) sb[i] = p.bias ? p.bias[i] : 0.f;
  // the xyz k-block of both A buffers: zero once (columns 3..63 stay zero, the first chunk is rewritten per tile)
  for (int i = threadIdx.x; i < 2 * 16384 / 16; i += GL_THREADS) {
    const int buf = i / (16384 / 16), off = i % (16384 / 16);
    *reinterpret_cast<uint4*>(sA + buf * a_bytes + (KB - 1) * 16384 + off * 16) = make_uint4(0u, 0u, 0u, 0u);
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::);
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ---------------- this CTA's half of the weight matrix, once
    if (elect_one()) {
      if (rank == 0) mbar_expect_tx(w_full, 2u * (uint32_t)(KB * w_block));
      for (int kb = 0; kb < KB; ++kb) tma_load_2d_pair(sW + kb * w_block, &tmW, w_full, kb * 64, rank * (p.N1 / 2));
    }
    __syncwarp();
  } else if (warp == 1) {
    if (rank == 0) {
      // ---------------- MMA issuer (leader CTA, for both SMs of the pair)
      const bool issuer = elect_one();
      const uint64_t dconst = umma_desc_sw128(0);
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.N1 >> 3) << 17) | ((uint32_t)((2 * TC_BM) >> 4) << 24);
      const uint32_t a_base = smem_u32(sA) >> 4, w_base = smem_u32(sW) >> 4;
      mbar_wait(w_full, 0);
      tc_fence_after();
      int it = 0;
      for (int tp = pair_id; tp < p.num_pairs; tp += pair_stride, ++it) {
        const int s = it & 1;
        const uint32_t par = (uint32_t)(it >> 1) & 1;
        mbar_wait(&a_full[s], par);
        mbar_wait(&acc_free[s], par ^ 1);
        tc_fence_after();
        if (issuer) {
          const uint32_t d = tmem_base + (uint32_t)(s * 256);
          for (int kb = 0; kb < KB; ++kb) {
            const uint64_t ad = dconst | (uint64_t)(a_base + ((s * a_bytes + kb * 16384) >> 4));
            const uint64_t bd = dconst | (uint64_t)(w_base + ((kb * w_block) >> 4));
#pragma unroll
            for (int k4 = 0; k4 < 4; ++k4) tc_mma_pair(d, ad + 2 * k4, bd + 2 * k4, idesc, (uint32_t)((kb | k4) != 0));
          }
          tc_commit_pair(&a_empty[s]);
          tc_commit_pair(&acc_full[s]);
        }
        __syncwarp();
      }
    }
  } else if (warp < 2 + GL_EPI_WARPS) {
    // ---------------- epilogue: accumulator -> bias, ReLU -> bf16 box -> TMA store; patch max read back from the box
    const int ew = warp - 2;
    const int q = warp & 3, h = ew >> 2;
    const uint32_t lane_field = (uint32_t)(q * 32) << 16;
    uint8_t* stg = sST + ew * 4096;
    const int ngroups = p.N1 / 64;
    int it = 0;
    for (int tp = pair_id; tp < p.num_pairs; tp += pair_stride, ++it) {
      const int s = it & 1;
      const int64_t row0 = (int64_t)(2 * tp + rank) * TC_BM + q * 32;
      const int nvalid = p.M - row0 >= 32 ? 32 : (p.M - row0 > 0 ? (int)(p.M - row0) : 0);
      mbar_wait(&acc_full[s], (uint32_t)(it >> 1) & 1);
      tc_fence_after();
      bool released = false;
      for (int gi = h; gi < ngroups; gi += 2) {
        const int n0 = gi * 64;
        const bool last = gi + 2 >= ngroups;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          float v[32];
          tc_ld32_issue(tmem_base + lane_field + (uint32_t)(s * 256 + n0 + half * 32), v);
          tc_ld_wait();
          if (last && half == 1) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cta(&acc_free[s], 0);
            released = true;
          }
          if (half == 0) {   // the previous box of this warp has left shared memory
            if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            __syncwarp();
          }
          const uint32_t sbias = smem_u32(sb + n0 + half * 32);
          const uint32_t rbase = smem_u32(stg) + lane * 128;
#pragma unroll
          for (int pc = 0; pc < 4; ++pc) {
            float4 b0, b1;
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(b0.x), "=f"(b0.y), "=f"(b0.z), "=f"(b0.w) : "r"(sbias + 32 * pc));
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(b1.x), "=f"(b1.y), "=f"(b1.z), "=f"(b1.w) : "r"(sbias + 32 * pc + 16));
            float* vv = &v[pc * 8];
            add2(vv[0], vv[1], vv[0], vv[1], b0.x, b0.y);
            add2(vv[2], vv[3], vv[2], vv[3], b0.z, b0.w);
            add2(vv[4], vv[5], vv[4], vv[5], b1.x, b1.y);
            add2(vv[6], vv[7], vv[6], vv[7], b1.z, b1.w);
            uint32_t pk[4];
            if (p.relu) {
              pk[0] = pack_bf16x2_relu(vv[0], vv[1]); pk[1] = pack_bf16x2_relu(vv[2], vv[3]);
              pk[2] = pack_bf16x2_relu(vv[4], vv[5]); pk[3] = pack_bf16x2_relu(vv[6], vv[7]);
            } else {
              pk[0] = pack_bf16x2(vv[0], vv[1]); pk[1] = pack_bf16x2(vv[2], vv[3]);
              pk[2] = pack_bf16x2(vv[4], vv[5]); pk[3] = pack_bf16x2(vv[6], vv[7]);
            }
            const uint32_t a = rbase + (((uint32_t)(pc + 4 * half) ^ (lane & 7)) << 4);
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(pk[0]), "r"(pk[1]), "r"(pk[2]), "r"(pk[3]) : "memory");
          }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0 && row0 < p.M) {
          asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                           reinterpret_cast<uint64_t>(&tmC)),
                       "r"(smem_u32(stg)), "r"(n0), "r"((int)row0)
                       : "memory");
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
        if (p.out_max_bf16 && row0 < p.M) {
          const uint32_t m = box_rows_max_bf16x2(smem_u32(stg), lane, nvalid);
          *reinterpret_cast<uint32_t*>(p.out_max_bf16 + (size_t)(row0 >> 5) * p.N1 + n0 + 2 * lane) = m;
        }
        __syncwarp();
      }
      if (!released) {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cta(&acc_free[s], 0);
      }
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  } else {
    // ---------------- gather producers: 16 rows per warp and tile, a half-warp per row and feature k-block
    const int pw = warp - 2 - GL_EPI_WARPS;
    const int hw = lane >> 4, hl = lane & 15;          // half-warp, lane inside it
    const IdxT* knn = reinterpret_cast<const IdxT*>(p.knn);
    int it = 0;
    for (int tp = pair_id; tp < p.num_pairs; tp += pair_stride, ++it) {
      const int s = it & 1;
      const int64_t trow0 = (int64_t)(2 * tp + rank) * TC_BM + pw * 16;     // first of this warp's 16 rows
      // source point of row trow0 + j (lanes 0..15 hold j = lane; rows beyond M re-read the last row, their results are clipped)
      int64_t src = 0;
      {
        int64_t r = trow0 + hl;
        if (r >= p.M) r = p.M - 1;
        const int64_t gq = p.g_begin + r / p.k;        // global group
        const int64_t b = gq / p.G;
        src = b * p.N + (int64_t)knn[gq * p.k + (r % p.k)];
      }
      // loads of all 16 rows x DB feature blocks in flight: 8 row pairs x DB float4 per lane
      float4 f[8][DB];
#pragma unroll
      for (int pr = 0; pr < 8; ++pr) {
        const int64_t sp = __shfl_sync(0xffffffffu, src, 2 * pr + hw);
        const float4* row = reinterpret_cast<const float4*>(p.feats + sp * p.D);
#pragma unroll
        for (int kb = 0; kb < DB; ++kb) f[pr][kb] = __ldg(row + kb * 16 + hl);
      }
      float px = 0.f, py = 0.f, pz = 0.f;
      if (lane < 16) {
        const float* pp = p.pts + src * 3;
        px = pp[0]; py = pp[1]; pz = pp[2];
      }
      mbar_wait(&a_empty[s], ((uint32_t)(it >> 1) & 1) ^ 1);      // the MMAs that read this buffer two tiles ago are done
      uint8_t* A = sA + s * a_bytes;
#pragma unroll
      for (int pr = 0; pr < 8; ++pr) {
        const int trow = pw * 16 + 2 * pr + hw;                    // row inside this CTA's 128-row tile
        const uint32_t rbase = smem_u32(A) + (uint32_t)trow * 128u + ((((uint32_t)hl >> 1) ^ (uint32_t)(trow & 7)) << 4) + (uint32_t)(hl & 1) * 8u;
#pragma unroll
        for (int kb = 0; kb < DB; ++kb) {
          const uint32_t lo = pack_bf16x2(f[pr][kb].x, f[pr][kb].y), hi = pack_bf16x2(f[pr][kb].z, f[pr][kb].w);
          asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(rbase + (uint32_t)kb * 16384u), "r"(lo), "r"(hi) : "memory");
        }
      }
      if (lane < 16) {   // xyz block: [x, y, z, 0 ...] in the row's first chunk (the rest of the row was zeroed once)
        const int trow = pw * 16 + lane;
        const uint32_t a = smem_u32(A) + (uint32_t)DB * 16384u + (uint32_t)trow * 128u + (((uint32_t)(trow & 7)) << 4);
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(pack_bf16x2(px, py)), "r"(pack_bf16x2(pz, 0.f)), "r"(0u), "r"(0u)
                     : "memory");
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the MMA (async proxy)
      __syncwarp();
      if (lane == 0) mbar_arrive_cta(&a_full[s], 0);
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
  }
}

static int gather_smem_bytes(int KB, int N1) {
  return KB * (N1 / 2) * 128 + 2 * KB * 16384 + GL_EPI_WARPS * 4096 + 1024 + 16 * 8 + 64 + 1024;
}

bool tc_gather_linear_supported(const p3tok_rows* R, int N1, int64_t k) {
  if (R->kind != 1 || (R->D != 64 && R->D != 128 && R->D != 256)) return false;
  if (N1 % 64 || N1 < 64 || N1 > 256 || k % 32) return false;
  if ((reinterpret_cast<uintptr_t>(R->feats) & 15) != 0) return false;
  return gather_smem_bytes(R->D / 64 + 1, N1) <= 227 * 1024;
}

// h = relu(W [feats | xyz | 0] + bias) for the rows of groups [g_begin, g_begin + rows / k); W: [N1, (D/64 + 1) * 64] bf16 in the
// rotated, padded column order (pad_weight_kernel, rot = 3).  out_bf16 [rows, N1], out_max_bf16 [rows / 32, N1].
int tc_gather_linear(const p3tok_rows* R, int64_t g_begin, int64_t rows, const __nv_bfloat16* W, int N1, const float* bias, int relu,
                     __nv_bfloat16* out_bf16, __nv_bfloat16* out_max_bf16, cudaStream_t s) {
  P3_REQUIRE(tc_gather_linear_supported(R, N1, R->k), P3TOK_ERR_UNSUPPORTED, "tc_gather_linear: unsupported shape D=%d N1=%d", R->D, N1);
  P3_REQUIRE(rows < (1ll << 31) - 512 && out_bf16, P3TOK_ERR_INVALID, "tc_gather_linear: bad arguments");
  if (rows == 0) return P3TOK_OK;
  GatherParams p;
  p.M = rows; p.g_begin = g_begin; p.N1 = N1; p.D = R->D; p.KB = R->D / 64 + 1;
  const int num_m_tiles = (int)((rows + TC_BM - 1) / TC_BM);
  p.num_pairs = (num_m_tiles + 1) / 2;
  p.N = R->N; p.G = R->G; p.k = R->k;
  p.pts = R->x; p.feats = R->feats; p.knn = R->knn_idx; p.idx64 = R->idx_dtype == P3TOK_I64;
  p.bias = bias; p.relu = relu; p.out_max_bf16 = out_max_bf16;
  CUtensorMap tw, tc;
  int rc = make_map(&tw, W, N1, (int64_t)p.KB * 64, N1 / 2);
  if (rc) return rc;
  rc = make_map(&tc, out_bf16, rows, N1, 32);
  if (rc) return rc;
  static thread_local bool configured[32] = {false};
  int dev = 0;
  P3_CUDA(cudaGetDevice(&dev));
  if (dev < 32 && !configured[dev]) {
#define GL_CFG(I, B) P3_CUDA(cudaFuncSetAttribute(tc_gather_linear_kernel<I, B>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024))
    GL_CFG(int32_t, 1); GL_CFG(int32_t, 2); GL_CFG(int32_t, 4); GL_CFG(int64_t, 1); GL_CFG(int64_t, 2); GL_CFG(int64_t, 4);
#undef GL_CFG
    configured[dev] = true;
  }
  const int max_pairs = num_sms() / 2;
  const int pairs = p.num_pairs < max_pairs ? p.num_pairs : max_pairs;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(pairs * 2));
  cfg.blockDim = dim3(GL_THREADS);
  cfg.dynamicSmemBytes = (size_t)gather_smem_bytes(p.KB, N1);
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
#define GL_LAUNCH(I, B) P3_CUDA(cudaLaunchKernelEx(&cfg, tc_gather_linear_kernel<I, B>, tw, tc, p))
  const int DB = R->D / 64;
  if (p.idx64) { if (DB == 1) GL_LAUNCH(int64_t, 1); else if (DB == 2) GL_LAUNCH(int64_t, 2); else GL_LAUNCH(int64_t, 4); }
  else         { if (DB == 1) GL_LAUNCH(int32_t, 1); else if (DB == 2) GL_LAUNCH(int32_t, 2); else GL_LAUNCH(int32_t, 4); }
#undef GL_LAUNCH
  count_launch();
  return P3TOK_OK;
}

}  // namespace p3tok

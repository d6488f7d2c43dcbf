// fps_culled.cu - farthest point sampling on a spatially sorted cloud: exact, but only the 32-point blocks whose running
// distances CAN change are touched per iteration.
//
// Replaces furthest_point_sample (reference src/data/sampler.py:4-30) / farthest_point_sampling
// (src/models/pix4point.py:8-53) for clouds of at most 8192 points, on the workspace p3tok_knn_prepare already builds for
// the kNN query of the same clouds (csrc/knn.cu: points sorted along a Z-order curve, 32 consecutive sorted points = one
// block with an exact bounding box).  Same picks as fps_kernel (csrc/fps.cu) bit for bit - same distance arithmetic
// ((dx*dx)+(dy*dy))+(dz*dz) with every operation rounded, same lowest-ORIGINAL-index tie-break - because a skipped block
// is one whose running distances provably do not change:
//   a point p of block b has  |p.x - c.x| >= gap_x = max(lo.x - c.x, c.x - hi.x, 0)  per axis, and fp32 subtraction,
//   multiplication and addition are monotonic, so evaluating the reference's own operation sequence on the per-axis gaps
//   gives a lower bound  lb <= d(p, c)  of the COMPUTED distance of every point in the box; if lb >= max over the block of
//   the running distance, fminf(md[p], d(p, c)) == md[p] for all its points and the block's maximum stays what it is.
// After ~100 centres a new centre changes ~13 of 256 blocks at N = 8192 (5 of 32 at N = 1024), so an iteration is a box
// test per block, a handful of block updates and an argmax over the BLOCK maxima - not a pass over the cloud (fps_kernel:
// 1180 cycles per iteration at N = 8192, 2 waves of 256 one-CTA clouds at BASELINE configs[2]).
//
// One CTA per cloud, W <= 8 warps (256 threads, <= 128 registers: two clouds per SM); block b belongs to warp b % W, slot b / W (a new centre changes spatially ADJACENT
// blocks - adjacent on the curve - so interleaving spreads the work of an iteration over the warps; the first version
// gave each warp consecutive blocks and one warp did all the updates: 3150 cycles per iteration).  Two views of a warp:
//   point view  lane l = point l of slot j: its running distance md[j] and original index live in REGISTERS (the slot loop
//               is fully unrolled, so the indices are static); coordinates come from shared memory (SoA, conflict-free);
//   slot view   lane j = slot j: the block's bounding box and the block's maximum running distance (float bits).
// Iteration: box tests (one per lane) -> ballot -> for the flagged slots: distance, fminf and ONE redux.sync.max for the
// new block maximum (independent across slots: their latencies overlap, the results are committed after the loop) ->
// thread-local maximum over the slots, argmax resolved only by the lanes that tie with the warp maximum (lowest original
// index), two redux.sync -> one __syncthreads and an argmax over the W warp candidates, as in fps_kernel.
// Shared memory: 12 B per point, so two 8192-point clouds fit one SM and BASELINE configs[2]'s 256 clouds are ONE wave.
// Measured (B200): worth it from N = 4096 up - below that an FPS iteration is the latency of the reduction chain, not the
// distance pass, and the sweep kernel is as fast (p3tok/ops.py dispatches).
#include <stdlib.h>

#include "common.cuh"

namespace p3tok {

constexpr int FPC_MAX_N = 8192;
constexpr int FPC_MAX_W = 8;

__device__ __forceinline__ void fpc_argmax(uint32_t& key, uint32_t& idx) {     // max key, lowest idx among the maxima
  const uint32_t m = __reduce_max_sync(0xffffffffu, key);
  const uint32_t cand = (key == m) ? idx : 0xffffffffu;
  idx = __reduce_min_sync(0xffffffffu, cand);
  key = m;
}

struct FpcCand {
  uint32_t key, idx, pos, pad;
};

template <int S>
__global__ void __launch_bounds__(32 * FPC_MAX_W, 2)
fps_culled_kernel(const float4* __restrict__ pts, const int* __restrict__ ids, const float* __restrict__ bb, int N, int nblk,
                  const int64_t* __restrict__ start_idx, int G, int64_t* __restrict__ out_idx) {
  extern __shared__ __align__(16) unsigned char fpc_smem[];
  const int npad = nblk * 32;
  float* sx = reinterpret_cast<float*>(fpc_smem);
  float* sy = sx + npad;
  float* sz = sy + npad;
  __shared__ uint32_t wkey[2][FPC_MAX_W];
  __shared__ unsigned long long gbest[2];       // (lowest original index << 32 | sorted position) among the points at the maximum
  __shared__ int s_start_pos;

  const int cloud = blockIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, W = blockDim.x >> 5;
  const float4* P = pts + (size_t)cloud * npad;
  const int* I = ids + (size_t)cloud * npad;
  const float* BB = bb + (size_t)cloud * nblk * 8;

  for (int i = threadIdx.x; i < npad; i += blockDim.x) {
    const float4 p = P[i];
    sx[i] = p.x; sy[i] = p.y; sz[i] = p.z;
  }
  int start = (int)start_idx[cloud];
  start = min(max(start, 0), N - 1);          // the C ABI cannot validate device data; an out-of-range start index is clamped

  // point view: this lane's point of every slot (slot j of warp w = block w + j * W)
  float md[S];
  uint32_t oid[S];                            // original index (0xffffffff: padding)
#pragma unroll
  for (int j = 0; j < S; ++j) {
    const int b = warp + j * W;
    int id = -1;
    if (b < nblk) id = I[b * 32 + lane];
    oid[j] = id >= 0 ? (uint32_t)id : 0xffffffffu;
    md[j] = id >= 0 ? 1e10f : -2.f;           // sampler.py:19; padding is never the maximum and never updated (fminf keeps -2)
    if (id == start) s_start_pos = b * 32 + lane;
  }
  // slot view: lane j <-> block warp + j * W
  const int myb = warp + lane * W;
  const bool slot_ok = lane < S && myb < nblk;
  float lox = 0.f, loy = 0.f, loz = 0.f, hix = 0.f, hiy = 0.f, hiz = 0.f;
  if (slot_ok) {
    lox = BB[myb * 8 + 0]; loy = BB[myb * 8 + 1]; loz = BB[myb * 8 + 2];
    hix = BB[myb * 8 + 3]; hiy = BB[myb * 8 + 4]; hiz = BB[myb * 8 + 5];
  }
  uint32_t bkey = slot_ok ? __float_as_uint(1e10f) : 0u;   // the block's maximum running distance (every block holds a real point)
  if (threadIdx.x == 0) { gbest[0] = 0xffffffffffffffffull; gbest[1] = 0xffffffffffffffffull; }
  __syncthreads();
  int far = start;
  int fpos = s_start_pos;
  float cx = sx[fpos], cy = sy[fpos], cz = sz[fpos];

  for (int g = 0; g < G; ++g) {
    if (threadIdx.x == 0) out_idx[(size_t)cloud * G + g] = far;
    if (g == G - 1) break;
    // ---- slot view: which of my warp's blocks can change?  lb = the reference's distance arithmetic on the per-axis gaps
    const float ax = fmaxf(fmaxf(__fsub_rn(lox, cx), __fsub_rn(cx, hix)), 0.f);
    const float ay = fmaxf(fmaxf(__fsub_rn(loy, cy), __fsub_rn(cy, hiy)), 0.f);
    const float az = fmaxf(fmaxf(__fsub_rn(loz, cz), __fsub_rn(cz, hiz)), 0.f);
    const float lb = __fadd_rn(__fadd_rn(__fmul_rn(ax, ax), __fmul_rn(ay, ay)), __fmul_rn(az, az));
    const bool need = slot_ok && !(lb >= __uint_as_float(bkey));
    const uint32_t mask = __ballot_sync(0xffffffffu, need);
    // ---- point view: update the flagged slots.  The slot bodies are static code (md[j], oid[j] in registers) reached
    // through a jump table, so an iteration pays only for the slots it touches (an unrolled "if (mask & bit)" ladder cost
    // ~3 instructions per slot per iteration, 460 instructions per warp per iteration at S = 32: issue-bound, slower than
    // the sweep).
    uint32_t mm = mask;
    while (mm) {
      const int j = __ffs(mm) - 1;
      mm &= mm - 1;
      const int pos = (warp + j * W) * 32 + lane;
      const float dx = __fsub_rn(sx[pos], cx), dy = __fsub_rn(sy[pos], cy), dz = __fsub_rn(sz[pos], cz);
      const float d = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
      float v = 0.f;
      switch (j) {
#define FPC_CASE(J) case J: if (J < S) { md[J < S ? J : 0] = fminf(md[J < S ? J : 0], d); v = md[J < S ? J : 0]; } break;
        FPC_CASE(0) FPC_CASE(1) FPC_CASE(2) FPC_CASE(3) FPC_CASE(4) FPC_CASE(5) FPC_CASE(6) FPC_CASE(7)
        FPC_CASE(8) FPC_CASE(9) FPC_CASE(10) FPC_CASE(11) FPC_CASE(12) FPC_CASE(13) FPC_CASE(14) FPC_CASE(15)
        FPC_CASE(16) FPC_CASE(17) FPC_CASE(18) FPC_CASE(19) FPC_CASE(20) FPC_CASE(21) FPC_CASE(22) FPC_CASE(23)
        FPC_CASE(24) FPC_CASE(25) FPC_CASE(26) FPC_CASE(27) FPC_CASE(28) FPC_CASE(29) FPC_CASE(30) FPC_CASE(31)
#undef FPC_CASE
        default: break;
      }
      const uint32_t m = __reduce_max_sync(0xffffffffu, v >= 0.f ? __float_as_uint(v) : 0u);
      if (lane == j) bkey = m;
    }
    // ---- argmax, level 1: maximum running distance of the cloud (thread-local over the slots, warp, CTA)
    float best = md[0];
#pragma unroll
    for (int j = 1; j < S; ++j) best = fmaxf(best, md[j]);
    const uint32_t mykey = best >= 0.f ? __float_as_uint(best) : 0u;
    uint32_t gmax = __reduce_max_sync(0xffffffffu, mykey);
    const int buf = g & 1;
    if (W > 1) {
      if (lane == 0) wkey[buf][warp] = gmax;
      __syncthreads();
      gmax = __reduce_max_sync(0xffffffffu, lane < W ? wkey[buf][lane] : 0u);
    }
    // ---- level 2: only lanes that hold the maximum resolve their lowest original index (one lane, absent exact ties)
    uint32_t idx = 0xffffffffu, pos = 0u;
    if (__any_sync(0xffffffffu, mykey == gmax)) {
      if (mykey == gmax) {
#pragma unroll
        for (int j = S - 1; j >= 0; --j)
          if (md[j] == best && oid[j] <= idx) { idx = oid[j]; pos = (uint32_t)((warp + j * W) * 32 + lane); }
      }
      const uint32_t widx = __reduce_min_sync(0xffffffffu, idx);
      if (idx == widx && idx != 0xffffffffu) atomicMin(&gbest[buf], ((unsigned long long)widx << 32) | pos);
    }
    if (W > 1) __syncthreads();
    else __syncwarp();
    {
      const unsigned long long w = gbest[buf];
      idx = (uint32_t)(w >> 32);
      pos = (uint32_t)w;
    }
    if (threadIdx.x == 0) gbest[buf ^ 1] = 0xffffffffffffffffull;      // re-armed one iteration ahead (a barrier lies between)
    far = (int)min(idx, (uint32_t)(N - 1));
    fpos = (int)min(pos, (uint32_t)(npad - 1));
    cx = sx[fpos]; cy = sy[fpos]; cz = sz[fpos];
  }
}

template <int S>
static int fpc_launch(const float4* pts, const int* ids, const float* bb, int B, int N, int nblk, const int64_t* start, int G,
                      int64_t* out, cudaStream_t s) {
  const int W = (nblk + S - 1) / S;
  const size_t smem = (size_t)nblk * 32 * 12;
  static thread_local bool configured[32] = {false};
  int dev = 0;
  P3_CUDA(cudaGetDevice(&dev));
  if (dev < 32 && !configured[dev]) {
    P3_CUDA(cudaFuncSetAttribute(fps_culled_kernel<S>, cudaFuncAttributeMaxDynamicSharedMemorySize, FPC_MAX_N * 12));
    configured[dev] = true;
  }
  fps_culled_kernel<S><<<(unsigned)B, 32 * W, smem, s>>>(pts, ids, bb, N, nblk, start, G, out);
  P3_LAUNCH_CHECK("fps_culled_kernel");
  return P3TOK_OK;
}

// knn.cu: the pieces of a prepared workspace
int kns_workspace_views(const void* workspace, int64_t workspace_bytes, int64_t B, int64_t N, const float4** pts, const int** ids,
                        const float** bb);

}  // namespace p3tok

using namespace p3tok;

extern "C" int p3tok_fps_sorted(const void* workspace, int64_t workspace_bytes, int64_t B, int64_t N, const int64_t* start_idx,
                                int64_t G, int64_t* out_idx, void* stream) {
  P3_REQUIRE(B >= 0 && N > 0 && G >= 0, P3TOK_ERR_INVALID, "fps_sorted: bad shape B=%lld N=%lld G=%lld", (long long)B, (long long)N,
             (long long)G);
  if (B == 0 || G == 0) return P3TOK_OK;
  P3_REQUIRE(workspace && start_idx && out_idx, P3TOK_ERR_INVALID, "fps_sorted: null pointer");
  P3_REQUIRE(N <= FPC_MAX_N, P3TOK_ERR_UNSUPPORTED, "fps_sorted: N=%lld exceeds %d (use p3tok_fps)", (long long)N, FPC_MAX_N);
  P3_REQUIRE(B < (1ll << 31) && B * G < (1ll << 40), P3TOK_ERR_UNSUPPORTED, "fps_sorted: batch too large");
  const float4* pts;
  const int* ids;
  const float* bb;
  int rc = kns_workspace_views(workspace, workspace_bytes, B, N, &pts, &ids, &bb);
  if (rc) return rc;
  const int nblk = (int)((N + 31) / 32);
  cudaStream_t s = as_stream(stream);
  // slots per warp: 8 blocks per warp up to 2048 points, 16 up to 4096, 32 up to 8192 (8 warps each)
  static int slots_env = -1;
  if (slots_env < 0) { const char* e = getenv("P3TOK_FPC_SLOTS"); slots_env = e ? atoi(e) : 0; }
  const int want = slots_env > 0 ? slots_env : (nblk <= 8 * FPC_MAX_W ? 8 : (nblk <= 16 * FPC_MAX_W ? 16 : 32));
  if (want <= 8 && nblk <= 8 * FPC_MAX_W) return fpc_launch<8>(pts, ids, bb, (int)B, (int)N, nblk, start_idx, (int)G, out_idx, s);
  if (want <= 16 && nblk <= 16 * FPC_MAX_W) return fpc_launch<16>(pts, ids, bb, (int)B, (int)N, nblk, start_idx, (int)G, out_idx, s);
  return fpc_launch<32>(pts, ids, bb, (int)B, (int)N, nblk, start_idx, (int)G, out_idx, s);
}

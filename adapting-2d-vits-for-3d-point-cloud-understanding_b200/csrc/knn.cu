// knn.cu - k nearest neighbours of each patch centre: register-tiled distance sweep with a
// warp-level top-k.
//
// Replaces _square_distance + knn_point (reference src/data/sampler.py:47-75; mode APF_SQ) and
// the cdist + topk inside group_knn (src/models/pix4point.py:79-89; mode P4P_CDIST).  The
// (B,G,N) distance matrix the reference materialises is never formed.
//
// One warp owns CPW centres.  The CTA streams the cloud through shared memory in tiles of
// float4 {x, y, z, |p|^2}; each lane takes one point per step and evaluates it against the
// warp's CPW centres (the point is loaded once, the centres live in registers).  Distances
// follow the reference's exact fp32 operation order (see oracle/p3tok_oracle.c header):
//   APF : d = ((-2*dot)+|c|^2)+|p|^2 with dot = fma(cz,pz, fma(cy,py, cx*px))
//   P4P : t = fma(1,|p|^2, fma(|c|^2,1, fma(-2cz,pz, fma(-2cy,py, (-2cx)*px)))), d = sqrt(max(t,0))
// Selection: each centre keeps its current k best as 32*KPL sorted 64-bit keys
// (ordered-distance << 32 | index) spread over the warp's lanes.  A point is a candidate only
// if its distance is strictly below the current k-th distance (points arrive in ascending index
// order, so an equal distance with a larger index can never win); candidates are appended to a
// small shared-memory queue by ballot/popc and merged 32 at a time with a warp bitonic sort +
// bitonic merges.  The result is ascending by (distance, index): the canonical instance of
// torch.topk's implementation-defined tie order.
#include "common.cuh"

namespace p3tok {

constexpr int KNN_WARPS = 8;
constexpr int KNN_CPW = 4;                 // centres per warp (register tile)
constexpr int KNN_TILE = 2048;             // points per shared-memory tile (32 KB)
constexpr uint64_t KNN_SENTINEL = 0xffffffffffffffffull;

__device__ __forceinline__ uint64_t shfl_xor_u64(uint64_t v, int m) {
  uint32_t lo = (uint32_t)v, hi = (uint32_t)(v >> 32);
  lo = __shfl_xor_sync(0xffffffffu, lo, m);
  hi = __shfl_xor_sync(0xffffffffu, hi, m);
  return ((uint64_t)hi << 32) | lo;
}
__device__ __forceinline__ uint64_t shfl_u64(uint64_t v, int src) {
  uint32_t lo = (uint32_t)v, hi = (uint32_t)(v >> 32);
  lo = __shfl_sync(0xffffffffu, lo, src);
  hi = __shfl_sync(0xffffffffu, hi, src);
  return ((uint64_t)hi << 32) | lo;
}
__device__ __forceinline__ uint64_t umin64(uint64_t a, uint64_t b) { return a < b ? a : b; }
__device__ __forceinline__ uint64_t umax64(uint64_t a, uint64_t b) { return a < b ? b : a; }

// ascending bitonic sort of one key per lane
__device__ __forceinline__ uint64_t warp_sort_asc(uint64_t v, int lane) {
#pragma unroll
  for (int k = 2; k <= 32; k <<= 1) {
#pragma unroll
    for (int j = k >> 1; j > 0; j >>= 1) {
      const uint64_t o = shfl_xor_u64(v, j);
      const bool up = ((lane & k) == 0);          // this k-block sorts ascending
      const bool lower = ((lane & j) == 0);
      v = (up == lower) ? umin64(v, o) : umax64(v, o);
    }
  }
  return v;
}
// input: bitonic sequence over lanes; output ascending
__device__ __forceinline__ uint64_t warp_bitonic_merge_asc(uint64_t v, int lane) {
#pragma unroll
  for (int j = 16; j > 0; j >>= 1) {
    const uint64_t o = shfl_xor_u64(v, j);
    v = ((lane & j) == 0) ? umin64(v, o) : umax64(v, o);
  }
  return v;
}

// Merge an ascending run `q` (one key per lane) into the ascending list L[0..KPL) (element
// j*32+lane), keeping the 32*KPL smallest.
template <int KPL>
__device__ __forceinline__ void list_insert_run(uint64_t (&L)[KPL], uint64_t q, int lane) {
#pragma unroll
  for (int j = 0; j < KPL; ++j) {
    const uint64_t r = shfl_u64(q, 31 - lane);     // reversed run -> L[j] ++ r is bitonic per lane pair
    const uint64_t lo = umin64(L[j], r);
    if (j + 1 < KPL) {
      const uint64_t hi = umax64(L[j], r);
      q = warp_bitonic_merge_asc(hi, lane);        // carry: the 32 largest, ascending
    }
    L[j] = warp_bitonic_merge_asc(lo, lane);
  }
}

template <int KPL, int MODE>
__global__ void __launch_bounds__(KNN_WARPS * 32)
knn_kernel(const float* __restrict__ x, int N, int pt_stride, const float* __restrict__ centres, int G,
           int k, void* __restrict__ idx_out, int idx_is_i64, float* __restrict__ dist_out) {
  __shared__ float4 tile[KNN_TILE];
  __shared__ uint64_t queue[KNN_WARPS * KNN_CPW][64];

  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const int b = blockIdx.y;
  const int g0 = (blockIdx.x * KNN_WARPS + warp) * KNN_CPW;   // first centre of this warp
  const float* P = x + (size_t)b * N * pt_stride;

  // centre registers (warp-uniform)
  float c0[KNN_CPW], c1[KNN_CPW], c2[KNN_CPW], cn[KNN_CPW];
#pragma unroll
  for (int c = 0; c < KNN_CPW; ++c) {
    const int g = min(g0 + c, G - 1);
    const float* cp = centres + ((size_t)b * G + g) * 3;
    const float cx = cp[0], cy = cp[1], cz = cp[2];
    cn[c] = sq3(cx, cy, cz);
    if (MODE == P3TOK_KNN_APF_SQ) {
      c0[c] = cx; c1[c] = cy; c2[c] = cz;
    } else {
      c0[c] = __fmul_rn(-2.f, cx); c1[c] = __fmul_rn(-2.f, cy); c2[c] = __fmul_rn(-2.f, cz);
    }
  }

  uint64_t L[KNN_CPW][KPL];
  float thr[KNN_CPW];     // candidates need d < thr  (or, P4P: t <= thr as a cheap pre-filter)
  float thr_d[KNN_CPW];   // exact current k-th distance (P4P), +inf while the list is not full
  int qn[KNN_CPW];
#pragma unroll
  for (int c = 0; c < KNN_CPW; ++c) {
#pragma unroll
    for (int j = 0; j < KPL; ++j) L[c][j] = KNN_SENTINEL;
    thr[c] = __int_as_float(0x7f800000);
    thr_d[c] = __int_as_float(0x7f800000);
    qn[c] = 0;
  }
  const int kth_lane = (k - 1) & 31, kth_j = (k - 1) >> 5;

  auto merge_queue = [&](int c, uint64_t (&Lc)[KPL], int cnt) {
    uint64_t* q = queue[warp * KNN_CPW + c];
    uint64_t v = (lane < cnt) ? q[lane] : KNN_SENTINEL;
    v = warp_sort_asc(v, lane);
    list_insert_run<KPL>(Lc, v, lane);
    // new threshold = k-th smallest so far
    uint64_t kth = 0;
#pragma unroll
    for (int j = 0; j < KPL; ++j)
      if (j == kth_j) kth = shfl_u64(Lc[j], kth_lane);
    const uint32_t ko = (uint32_t)(kth >> 32);
    const float kd = (ko == 0xffffffffu) ? __int_as_float(0x7f800000) : ord2f(ko);
    return kd;
  };

  for (int base = 0; base < N; base += KNN_TILE) {
    __syncthreads();
    const int cnt = min(KNN_TILE, N - base);
    for (int i = t; i < cnt; i += KNN_WARPS * 32) {
      const float* pp = P + (size_t)(base + i) * pt_stride;
      const float px = pp[0], py = pp[1], pz = pp[2];
      tile[i] = make_float4(px, py, pz, sq3(px, py, pz));
    }
    __syncthreads();
    if (g0 >= G) continue;   // warp without centres still helps loading tiles
    for (int i0 = 0; i0 < cnt; i0 += 32) {
      const int i = i0 + lane;
      const bool valid = i < cnt;
      const float4 p = tile[valid ? i : 0];
      // all CPW distance chains first (independent: instruction-level parallelism), ONE vote for the common case
      // "no candidate for any of the warp's centres", then the per-centre handling
      float dc[KNN_CPW];
      bool pc[KNN_CPW];
      bool any = false;
#pragma unroll
      for (int c = 0; c < KNN_CPW; ++c) {
        if (MODE == P3TOK_KNN_APF_SQ) {
          const float dot = __fmaf_rn(c2[c], p.z, __fmaf_rn(c1[c], p.y, __fmul_rn(c0[c], p.x)));
          float tt = __fmul_rn(-2.f, dot);
          tt = __fadd_rn(tt, cn[c]);
          dc[c] = __fadd_rn(__fadd_rn(tt, p.w), 0.f);
          pc[c] = valid && (dc[c] < thr[c]);
        } else {
          float tt = __fmul_rn(c0[c], p.x);
          tt = __fmaf_rn(c1[c], p.y, tt);
          tt = __fmaf_rn(c2[c], p.z, tt);
          tt = __fadd_rn(tt, cn[c]);      // fma(|c|^2, 1, t)
          tt = __fadd_rn(tt, p.w);        // fma(1, |p|^2, t)
          dc[c] = tt;
          pc[c] = valid && (tt <= thr[c]); // conservative: exact sqrt test below
        }
        any = any || pc[c];
      }
      if (!__any_sync(0xffffffffu, any)) continue;
#pragma unroll
      for (int c = 0; c < KNN_CPW; ++c) {
        float d = dc[c];
        bool pass = pc[c];
        uint32_t ball = __ballot_sync(0xffffffffu, pass);
        if (ball == 0) continue;
        if (MODE == P3TOK_KNN_P4P_CDIST) {
          d = __fadd_rn(__fsqrt_rn(fmaxf(d, 0.f)), 0.f);
          pass = pass && (d < thr_d[c]);
          ball = __ballot_sync(0xffffffffu, pass);
          if (ball == 0) continue;
        }
        uint64_t* q = queue[warp * KNN_CPW + c];
        if (pass) {
          const int pos = qn[c] + __popc(ball & ((1u << lane) - 1u));
          q[pos] = ((uint64_t)f2ord(d) << 32) | (uint32_t)(base + i);
        }
        qn[c] += __popc(ball);
        __syncwarp();
        if (qn[c] >= 32) {
          const float kd = merge_queue(c, L[c], 32);
          __syncwarp();
          if (lane + 32 < qn[c]) q[lane] = q[lane + 32];
          qn[c] -= 32;
          __syncwarp();
          if (MODE == P3TOK_KNN_APF_SQ) {
            thr[c] = kd;
          } else {
            thr_d[c] = kd;
            // t > kd*kd*(1+2^-22) (rounded up) implies sqrt_rn(t) >= kd: safe to drop early
            thr[c] = (kd == __int_as_float(0x7f800000)) ? kd
                     : __fmul_ru(__fmul_ru(kd, kd), 1.0000002384185791f);
          }
        }
      }
    }
  }
  if (g0 >= G) return;
#pragma unroll
  for (int c = 0; c < KNN_CPW; ++c) {
    if (qn[c] > 0) (void)merge_queue(c, L[c], qn[c]);
    const int g = g0 + c;
    if (g >= G) continue;
    const size_t o = ((size_t)b * G + g) * k;
#pragma unroll
    for (int j = 0; j < KPL; ++j) {
      const int r = j * 32 + lane;
      if (r < k) {
        const uint32_t id = (uint32_t)L[c][j];
        if (idx_is_i64) reinterpret_cast<int64_t*>(idx_out)[o + r] = (int64_t)id;
        else reinterpret_cast<int32_t*>(idx_out)[o + r] = (int32_t)id;
        if (dist_out) dist_out[o + r] = ord2f((uint32_t)(L[c][j] >> 32));
      }
    }
  }
}

template <int KPL>
static int knn_launch(const float* x, int B, int N, int pt_stride, const float* centres, int G, int k,
                      int mode, void* idx_out, int i64, float* dist_out, cudaStream_t s) {
  dim3 grid((unsigned)((G + KNN_WARPS * KNN_CPW - 1) / (KNN_WARPS * KNN_CPW)), (unsigned)B);
  if (mode == P3TOK_KNN_APF_SQ)
    knn_kernel<KPL, P3TOK_KNN_APF_SQ><<<grid, KNN_WARPS * 32, 0, s>>>(x, N, pt_stride, centres, G, k, idx_out, i64, dist_out);
  else
    knn_kernel<KPL, P3TOK_KNN_P4P_CDIST><<<grid, KNN_WARPS * 32, 0, s>>>(x, N, pt_stride, centres, G, k, idx_out, i64, dist_out);
  P3_LAUNCH_CHECK("knn_kernel");
  return P3TOK_OK;
}

}  // namespace p3tok

using namespace p3tok;

extern "C" int p3tok_knn(const float* x, int64_t B, int64_t N, int64_t pt_stride, const float* centres,
                         int64_t G, int64_t k, int mode, void* idx_out, int idx_dtype, float* dist_out,
                         void* stream) {
  P3_REQUIRE(B >= 0 && N > 0 && G >= 0 && pt_stride >= 3, P3TOK_ERR_INVALID, "knn: bad shape");
  P3_REQUIRE(mode == P3TOK_KNN_APF_SQ || mode == P3TOK_KNN_P4P_CDIST, P3TOK_ERR_INVALID, "knn: bad mode %d", mode);
  P3_REQUIRE(idx_dtype == P3TOK_I64 || idx_dtype == P3TOK_I32, P3TOK_ERR_INVALID, "knn: idx dtype must be i32/i64");
  // torch.topk raises when k > N (sampler.py:74); mirror it as an error code
  P3_REQUIRE(k >= 1 && k <= N, P3TOK_ERR_INVALID, "knn: k=%lld out of range for N=%lld", (long long)k, (long long)N);
  P3_REQUIRE(k <= 128, P3TOK_ERR_UNSUPPORTED, "knn: k=%lld > 128", (long long)k);
  P3_REQUIRE(N < (1ll << 31) && B < 65536, P3TOK_ERR_UNSUPPORTED, "knn: N or B too large");
  if (B == 0 || G == 0) return P3TOK_OK;
  P3_REQUIRE(x && centres && idx_out, P3TOK_ERR_INVALID, "knn: null pointer");
  cudaStream_t s = as_stream(stream);
  const int i64 = idx_dtype == P3TOK_I64;
  if (k <= 32) return knn_launch<1>(x, (int)B, (int)N, (int)pt_stride, centres, (int)G, (int)k, mode, idx_out, i64, dist_out, s);
  if (k <= 64) return knn_launch<2>(x, (int)B, (int)N, (int)pt_stride, centres, (int)G, (int)k, mode, idx_out, i64, dist_out, s);
  return knn_launch<4>(x, (int)B, (int)N, (int)pt_stride, centres, (int)G, (int)k, mode, idx_out, i64, dist_out, s);
}

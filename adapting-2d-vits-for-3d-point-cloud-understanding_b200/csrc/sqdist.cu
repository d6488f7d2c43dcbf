// sqdist.cu - the materialised squared-distance matrix of the reference's _square_distance (src/data/sampler.py:47-62).
//
// The tokenizer never needs this matrix - p3tok_knn* fold the same arithmetic into the selection and the (B,S,N) tensor
// the reference writes (C2: 134 MB, C4: 8.6 GB) never exists.  The entry point is here for callers that use
// _square_distance on its own, and as a direct window onto the kNN kernels' distance arithmetic for the parity tests.
//
// Arithmetic (bit-exact restatement of  -2*matmul(src, dst^T) + |src|^2 + |dst|^2  on the reference's CPU path, verified in
// the build container, oracle/p3tok_oracle.c): dot = fma(sz,dz, fma(sy,dy, sx*dx));  d = ((-2*dot) + |s|^2) + |d|^2,
// |v|^2 = ((vx*vx)+(vy*vy))+(vz*vz); every operation individually rounded.  Values can be slightly negative (the reference's
// are: SURVEY.md 8a3); no clamp.
//
// Bound: the HBM write of the matrix (4 B per pair; the inputs are 12 B per point).  A CTA owns a tile of 1024 consecutive
// dst points of one cloud - four per thread, with their norms, in registers for the whole kernel - and walks the cloud's src
// rows (staged 256 at a time in shared memory with their norms, read back as one broadcast LDS.128 per row): per pair 7
// floating-point instructions, per row and warp one 512-byte contiguous store (128-bit per thread when N % 4 == 0).  The
// first version - one thread per pair with a 64-bit div / mod and three strided loads each - wrote 1.0-1.2 TB/s.
#include "common.cuh"

namespace p3tok {

constexpr int SQ_THREADS = 256;
constexpr int SQ_PPT = 4;                       // dst points per thread
constexpr int SQ_TILE = SQ_THREADS * SQ_PPT;    // dst points per CTA
constexpr int SQ_ROWS = 256;                    // src rows staged per pass

template <bool VEC>
__global__ void __launch_bounds__(SQ_THREADS)
sqdist_kernel(const float* __restrict__ src, const float* __restrict__ dst, int S, int N, int64_t dst_stride, int tiles,
              int rows_per_z, float* __restrict__ out) {
  __shared__ float4 cs[SQ_ROWS];
  const int tid = threadIdx.x;
  const int tile = blockIdx.x % tiles;
  const int64_t b = blockIdx.x / tiles;
  const int n0 = tile * SQ_TILE + tid * SQ_PPT;
  float px[SQ_PPT], py[SQ_PPT], pz[SQ_PPT], pn[SQ_PPT];
#pragma unroll
  for (int j = 0; j < SQ_PPT; ++j) {
    const bool ok = n0 + j < N;
    const float* p = dst + (b * N + (ok ? n0 + j : 0)) * dst_stride;
    px[j] = ok ? __ldg(p) : 0.f;
    py[j] = ok ? __ldg(p + 1) : 0.f;
    pz[j] = ok ? __ldg(p + 2) : 0.f;
    pn[j] = sq3(px[j], py[j], pz[j]);
  }
  const int s_begin = blockIdx.y * rows_per_z;
  const int s_end = min(S, s_begin + rows_per_z);
  for (int s0 = s_begin; s0 < s_end; s0 += SQ_ROWS) {
    const int cnt = min(SQ_ROWS, s_end - s0);
    __syncthreads();                             // the previous pass is done with cs
    for (int r = tid; r < cnt; r += SQ_THREADS) {
      const float* c = src + (b * S + s0 + r) * 3;
      const float cx = __ldg(c), cy = __ldg(c + 1), cz = __ldg(c + 2);
      cs[r] = make_float4(cx, cy, cz, sq3(cx, cy, cz));
    }
    __syncthreads();
    if (n0 >= N) continue;
    float* o = out + (b * S + s0) * (int64_t)N + n0;
    for (int r = 0; r < cnt; ++r, o += N) {
      const float4 c = cs[r];
      float d[SQ_PPT];
#pragma unroll
      for (int j = 0; j < SQ_PPT; ++j) {
        const float dot = __fmaf_rn(c.z, pz[j], __fmaf_rn(c.y, py[j], __fmul_rn(c.x, px[j])));
        d[j] = __fadd_rn(__fadd_rn(__fmul_rn(-2.f, dot), c.w), pn[j]);
      }
      if (VEC) {                                 // N % 4 == 0 and a 16-byte aligned matrix: the whole quad is in range
        *reinterpret_cast<float4*>(o) = make_float4(d[0], d[1], d[2], d[3]);
      } else {
#pragma unroll
        for (int j = 0; j < SQ_PPT; ++j)
          if (n0 + j < N) o[j] = d[j];
      }
    }
  }
}

}  // namespace p3tok

using namespace p3tok;

extern "C" int p3tok_square_distance(const float* src, int64_t B, int64_t S, const float* dst, int64_t N,
                                     int64_t dst_stride, float* out, void* stream) {
  P3_REQUIRE(B >= 0 && S >= 0 && N >= 0 && dst_stride >= 3, P3TOK_ERR_INVALID,
             "square_distance: bad shape B=%lld S=%lld N=%lld stride=%lld", (long long)B, (long long)S, (long long)N,
             (long long)dst_stride);
  if (B == 0 || S == 0 || N == 0) return P3TOK_OK;
  P3_REQUIRE(src && dst && out, P3TOK_ERR_INVALID, "square_distance: null pointer");
  P3_REQUIRE(S < (1ll << 31) && N < (1ll << 31) - SQ_TILE && S * N < (1ll << 40) / (B > 0 ? B : 1), P3TOK_ERR_UNSUPPORTED,
             "square_distance: %lld x %lld x %lld pairs (use p3tok_knn, which never materialises the matrix)", (long long)B,
             (long long)S, (long long)N);
  const int64_t tiles = (N + SQ_TILE - 1) / SQ_TILE;
  P3_REQUIRE(tiles * B < (1ll << 31), P3TOK_ERR_UNSUPPORTED, "square_distance: batch too large");
  // split the src rows over gridDim.y only while the (tile, cloud) grid alone leaves SMs idle
  int64_t zs = (148 * 8 + tiles * B - 1) / (tiles * B);
  const int64_t max_zs = (S + SQ_ROWS - 1) / SQ_ROWS;
  if (zs > max_zs) zs = max_zs;
  if (zs > 65535) zs = 65535;
  if (zs < 1) zs = 1;
  int64_t rows_per_z = (S + zs - 1) / zs;
  rows_per_z = (rows_per_z + SQ_ROWS - 1) / SQ_ROWS * SQ_ROWS;
  zs = (S + rows_per_z - 1) / rows_per_z;
  const dim3 grid((unsigned)(tiles * B), (unsigned)zs);
  const bool vec = N % 4 == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0;
  if (vec)
    sqdist_kernel<true><<<grid, SQ_THREADS, 0, as_stream(stream)>>>(src, dst, (int)S, (int)N, dst_stride, (int)tiles,
                                                                   (int)rows_per_z, out);
  else
    sqdist_kernel<false><<<grid, SQ_THREADS, 0, as_stream(stream)>>>(src, dst, (int)S, (int)N, dst_stride, (int)tiles,
                                                                    (int)rows_per_z, out);
  P3_LAUNCH_CHECK("sqdist_kernel");
  return P3TOK_OK;
}

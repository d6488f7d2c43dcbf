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
// One thread per output element, consecutive threads along the dst axis (coalesced stores); HBM-write bound: 4 B per pair.
#include "common.cuh"

namespace p3tok {

__global__ void __launch_bounds__(256)
sqdist_kernel(const float* __restrict__ src, const float* __restrict__ dst, int64_t S, int64_t N, int64_t dst_stride,
              int64_t total, float* __restrict__ out) {
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t n = e % N;
    const int64_t bs = e / N;            // b*S + s
    const int64_t b = bs / S;
    const float* c = src + bs * 3;
    const float* p = dst + (b * N + n) * dst_stride;
    const float cx = __ldg(c), cy = __ldg(c + 1), cz = __ldg(c + 2);
    const float px = __ldg(p), py = __ldg(p + 1), pz = __ldg(p + 2);
    const float dot = __fmaf_rn(cz, pz, __fmaf_rn(cy, py, __fmul_rn(cx, px)));
    float t = __fmul_rn(-2.f, dot);
    t = __fadd_rn(t, sq3(cx, cy, cz));
    out[e] = __fadd_rn(t, sq3(px, py, pz));
  }
}

}  // namespace p3tok

using namespace p3tok;

extern "C" int p3tok_square_distance(const float* src, int64_t B, int64_t S, const float* dst, int64_t N,
                                     int64_t dst_stride, float* out, void* stream) {
  P3_REQUIRE(B >= 0 && S >= 0 && N >= 0 && dst_stride >= 3, P3TOK_ERR_INVALID,
             "square_distance: bad shape B=%lld S=%lld N=%lld stride=%lld", (long long)B, (long long)S, (long long)N,
             (long long)dst_stride);
  const int64_t total = B * S * N;
  if (total == 0) return P3TOK_OK;
  P3_REQUIRE(src && dst && out, P3TOK_ERR_INVALID, "square_distance: null pointer");
  P3_REQUIRE(total < (1ll << 40), P3TOK_ERR_UNSUPPORTED, "square_distance: %lld pairs (use p3tok_knn, which never "
             "materialises the matrix)", (long long)total);
  int64_t blocks = (total + 255) / 256;
  if (blocks > 148 * 64) blocks = 148 * 64;   // grid-stride beyond 64 resident-CTA waves' worth of blocks
  sqdist_kernel<<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(src, dst, S, N, dst_stride, total, out);
  P3_LAUNCH_CHECK("sqdist_kernel");
  return P3TOK_OK;
}

// mlp_f32.cu - fp32 patch embedding on CUDA cores (the rtol-1e-4 precision mode).
//
// Replaces Encoder.get_features (reference src/models/apf.py:145-169) and the conv1/conv2/pool
// half of a P3Embed stage (src/models/pix4point.py:179-188) with eval-mode BatchNorm folded on
// the host.  Layer by layer: a 128x128x8 register-tiled SGEMM (8x8 outputs per thread, FFMA,
// fp32 accumulate, sequential-K order) with bias / per-group bias / ReLU fused in the epilogue,
// and a group max over the k rows of each patch.  The concat layer W.[g||f] is evaluated as
// W_g.g (once per group, becomes a per-group bias) + W_f.f, so the (rows, 2E) concat tensor is
// never built.  Activations ping-pong through a caller-provided workspace in chunks of groups.
// The tensor-core bf16 path (embed_tc.cu) is the throughput mode; this is the accuracy mode and
// its on-device reference.
#include "embed.cuh"

namespace p3tok {

constexpr int SG_BM = 128, SG_BN = 128, SG_BK = 8, SG_THREADS = 256;

__global__ void __launch_bounds__(SG_THREADS)
sgemm_bias_act_kernel(const float* __restrict__ A, int64_t M, int K, const float* __restrict__ W, int N,
                      const float* __restrict__ bias, const float* __restrict__ gbias, int rows_per_group,
                      int relu, float* __restrict__ C, int seg_rows, int seg_skip) {
  // relu: 0 none, 1 ReLU, 2 exact GELU (nn.GELU default).  Output row of input row m is m + (m / seg_rows + 1) *
  // seg_skip: with seg_skip = 1 every segment of seg_rows rows is preceded by one row left for the caller (the cls
  // token of pix4point.py:248-252); seg_skip = 0 is the plain layout.
  __shared__ __align__(16) float As[2][SG_BK][SG_BM];
  __shared__ __align__(16) float Ws[2][SG_BK][SG_BN];
  const int t = threadIdx.x;
  const int64_t m0 = (int64_t)blockIdx.x * SG_BM;
  const int n0 = blockIdx.y * SG_BN;
  const int tx = t & 15, ty = t >> 4;          // 16 x 16 threads, 8x8 outputs each
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  // loader mapping: element e = t + i*256 of the 128x8 tile: row = e >> 3, kk = e & 7
  float ra[4], rw[4];
  auto gload = [&](int k0) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int e = t + i * SG_THREADS;
      const int r = e >> 3, kk = e & 7;
      const int64_t m = m0 + r;
      const int n = n0 + r;
      const int kg = k0 + kk;
      ra[i] = (m < M && kg < K) ? A[m * K + kg] : 0.f;
      rw[i] = (n < N && kg < K) ? W[(int64_t)n * K + kg] : 0.f;
    }
  };
  auto sstore = [&](int buf) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int e = t + i * SG_THREADS;
      const int r = e >> 3, kk = e & 7;
      As[buf][kk][r] = ra[i];
      Ws[buf][kk][r] = rw[i];
    }
  };
  const int nk = (K + SG_BK - 1) / SG_BK;
  gload(0);
  sstore(0);
  __syncthreads();
  for (int kt = 0; kt < nk; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < nk) gload((kt + 1) * SG_BK);
#pragma unroll
    for (int kk = 0; kk < SG_BK; ++kk) {
      float a[8], w[8];
      *reinterpret_cast<float4*>(&a[0]) = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 4]);
      *reinterpret_cast<float4*>(&a[4]) = *reinterpret_cast<const float4*>(&As[buf][kk][64 + ty * 4]);
      *reinterpret_cast<float4*>(&w[0]) = *reinterpret_cast<const float4*>(&Ws[buf][kk][tx * 4]);
      *reinterpret_cast<float4*>(&w[4]) = *reinterpret_cast<const float4*>(&Ws[buf][kk][64 + tx * 4]);
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], w[j], acc[i][j]);
    }
    if (kt + 1 < nk) {
      sstore(buf ^ 1);
      __syncthreads();
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int64_t m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (m >= M) continue;
    const float* gb = gbias ? gbias + (m / rows_per_group) * N : nullptr;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int n = n0 + (j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4));
      if (n >= N) continue;
      float v = acc[i][j];
      if (bias) v += bias[n];
      if (gb) v += gb[n];
      if (relu == 1) v = fmaxf(v, 0.f);
      else if (relu == 2) v = 0.5f * v * (1.f + erff(v * 0.70710678118654752f));
      C[(m + (seg_skip ? (m / seg_rows + 1) * seg_skip : 0)) * N + n] = v;
    }
  }
}

__global__ void group_max_kernel(const float* __restrict__ in, int64_t ngroups, int k, int C, int relu,
                                 float* __restrict__ out) {
  const int64_t total = ngroups * C;
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total;
       e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t g = e / C;
    const int c = (int)(e - g * C);
    const float* p = in + g * k * C + c;
    float m = p[0];
    for (int r = 1; r < k; ++r) m = fmaxf(m, p[(int64_t)r * C]);
    out[e] = relu ? fmaxf(m, 0.f) : m;
  }
}

template <typename IdxT>
__global__ void build_rows_kernel(p3tok_rows R, int64_t g_begin, int64_t g_count, float* __restrict__ X) {
  const int cin = R.kind == 0 ? 2 * R.C : 3 + R.D;
  const int64_t total = g_count * R.k * cin;
  const IdxT* knn = reinterpret_cast<const IdxT*>(R.knn_idx);
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total;
       e += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(e % cin);
    const int64_t r = e / cin;
    const int n = (int)(r % R.k);
    const int64_t bj = g_begin + r / R.k;          // output group (b*G + j)
    const int64_t b = bj / R.G;
    const int64_t g = R.perm ? R.perm[bj] : (bj - b * R.G);
    const int64_t src = (b * R.G + g) * R.k + n;
    const int64_t ni = (int64_t)knn[src];
    float v;
    if (R.kind == 0) {
      const float* crow = R.x + (b * R.N + R.ctr_idx[b * R.G + g]) * R.C;
      v = c < R.C ? __fsub_rn(R.x[(b * R.N + ni) * R.C + c], crow[c]) : crow[c - R.C];
    } else {
      v = c < 3 ? R.x[(b * R.N + ni) * 3 + c] : R.feats[(b * R.N + ni) * R.D + (c - 3)];
    }
    X[e] = v;
  }
}

static inline unsigned grid_1d(int64_t total, int threads) {
  int64_t b = (total + threads - 1) / threads;
  const int64_t cap = 148 * 16;
  return (unsigned)(b < 1 ? 1 : (b > cap ? cap : b));
}

int build_rows_f32(const p3tok_rows* R, int64_t g_begin, int64_t g_count, float* X, cudaStream_t s) {
  const int cin = R->kind == 0 ? 2 * R->C : 3 + R->D;
  const int64_t total = g_count * R->k * cin;
  if (total == 0) return P3TOK_OK;
  if (R->idx_dtype == P3TOK_I64)
    build_rows_kernel<int64_t><<<grid_1d(total, 256), 256, 0, s>>>(*R, g_begin, g_count, X);
  else
    build_rows_kernel<int32_t><<<grid_1d(total, 256), 256, 0, s>>>(*R, g_begin, g_count, X);
  P3_LAUNCH_CHECK("build_rows_kernel");
  return P3TOK_OK;
}

static int linear_f32(const float* A, int64_t M, int64_t K, const float* W, int64_t N, const float* bias,
                      const float* gbias, int64_t rpg, int relu, float* C, cudaStream_t s, int seg_rows = 1,
                      int seg_skip = 0) {
  if (M == 0 || N == 0) return P3TOK_OK;
  dim3 grid((unsigned)((M + SG_BM - 1) / SG_BM), (unsigned)((N + SG_BN - 1) / SG_BN));
  sgemm_bias_act_kernel<<<grid, SG_THREADS, 0, s>>>(A, M, (int)K, W, (int)N, bias, gbias, (int)rpg, relu, C, seg_rows, seg_skip);
  P3_LAUNCH_CHECK("sgemm_bias_act_kernel");
  return P3TOK_OK;
}

static int group_max(const float* in, int64_t ng, int64_t k, int64_t C, int relu, float* out, cudaStream_t s) {
  if (ng * C == 0) return P3TOK_OK;
  group_max_kernel<<<grid_1d(ng * C, 256), 256, 0, s>>>(in, ng, (int)k, (int)C, relu, out);
  P3_LAUNCH_CHECK("group_max_kernel");
  return P3TOK_OK;
}

constexpr int64_t F32_CHUNK_GROUPS = 8192;

static int64_t max_width(const p3tok_mlp* m) {
  int64_t w = m->cin;
  for (int i = 0; i < m->n_pre; ++i) w = w > m->pre_dim[i] ? w : m->pre_dim[i];
  w = w > m->mid_dim ? w : m->mid_dim;
  w = w > m->out_dim ? w : m->out_dim;
  return w;
}

int64_t patch_embed_f32_workspace(const p3tok_mlp* m, int64_t ngroups, int64_t k) {
  const int64_t cg = ngroups < F32_CHUNK_GROUPS ? ngroups : F32_CHUNK_GROUPS;
  const int64_t rows = cg * k;
  const int64_t act = align_up(rows * max_width(m) * 4, 256);
  const int64_t F = m->pre_dim[m->n_pre - 1];
  return 2 * act + align_up(cg * F * 4, 256) + align_up(cg * m->mid_dim * 4, 256) + 256;
}

int patch_embed_f32(const p3tok_rows* R, const p3tok_mlp* m, void* ws, int64_t ws_bytes, float* tokens,
                    cudaStream_t s) {
  const int64_t ngroups = R->B * R->G, k = R->k;
  P3_REQUIRE(ws_bytes >= patch_embed_f32_workspace(m, ngroups, k), P3TOK_ERR_WORKSPACE,
             "patch_embed(f32): workspace %lld < %lld bytes", (long long)ws_bytes,
             (long long)patch_embed_f32_workspace(m, ngroups, k));
  const int64_t cg = ngroups < F32_CHUNK_GROUPS ? ngroups : F32_CHUNK_GROUPS;
  const int64_t act = align_up(cg * k * max_width(m) * 4, 256);
  const int64_t F = m->pre_dim[m->n_pre - 1];
  char* base = reinterpret_cast<char*>((reinterpret_cast<uintptr_t>(ws) + 255) / 256 * 256);
  float* buf0 = reinterpret_cast<float*>(base);
  float* buf1 = reinterpret_cast<float*>(base + act);
  float* gmax = reinterpret_cast<float*>(base + 2 * act);
  float* gbias = reinterpret_cast<float*>(base + 2 * act + align_up(cg * F * 4, 256));
  for (int64_t g0 = 0; g0 < ngroups; g0 += cg) {
    const int64_t gc = (ngroups - g0) < cg ? (ngroups - g0) : cg;
    const int64_t rows = gc * k;
    const float* h;
    float* nxt;
    if (R->kind == 2) {
      h = R->x + g0 * k * m->cin;
      nxt = buf0;
    } else {
      int rc = build_rows_f32(R, g0, gc, buf0, s);
      if (rc) return rc;
      h = buf0;
      nxt = buf1;
    }
    int64_t kin = m->cin;
    for (int i = 0; i < m->n_pre; ++i) {
      int rc = linear_f32(h, rows, kin, (const float*)m->w_pre[i], m->pre_dim[i], m->b_pre[i], nullptr, 1,
                          m->pre_relu[i], nxt, s);
      if (rc) return rc;
      h = nxt;
      nxt = (nxt == buf0) ? buf1 : buf0;
      kin = m->pre_dim[i];
    }
    int rc = group_max(h, gc, k, F, 0, gmax, s);
    if (rc) return rc;
    rc = linear_f32(gmax, gc, F, (const float*)m->w_mid_g, m->mid_dim, m->b_mid, nullptr, 1, 0, gbias, s);
    if (rc) return rc;
    rc = linear_f32(h, rows, F, (const float*)m->w_mid_f, m->mid_dim, nullptr, gbias, k, 1, nxt, s);
    if (rc) return rc;
    const float* h2 = nxt;
    float* o = (nxt == buf0) ? buf1 : buf0;
    rc = linear_f32(h2, rows, m->mid_dim, (const float*)m->w_out, m->out_dim, m->b_out, nullptr, 1, 0, o, s);
    if (rc) return rc;
    rc = group_max(o, gc, k, m->out_dim, m->out_relu, tokens + g0 * m->out_dim, s);
    if (rc) return rc;
  }
  return P3TOK_OK;
}

}  // namespace p3tok

using namespace p3tok;

extern "C" int p3tok_linear_f32(const float* A, int64_t M, int64_t K, const float* W, int64_t N, const float* bias,
                                const float* gbias, int64_t rows_per_group, int relu, float* C, void* stream) {
  P3_REQUIRE(M >= 0 && K > 0 && N > 0 && K < (1 << 30) && N < (1 << 30), P3TOK_ERR_INVALID, "linear_f32: bad shape");
  P3_REQUIRE(!gbias || rows_per_group > 0, P3TOK_ERR_INVALID, "linear_f32: rows_per_group must be > 0");
  if (M == 0) return P3TOK_OK;
  P3_REQUIRE(A && W && C, P3TOK_ERR_INVALID, "linear_f32: null pointer");
  return linear_f32(A, M, K, W, N, bias, gbias, rows_per_group > 0 ? rows_per_group : 1, relu, C, as_stream(stream));
}

extern "C" int p3tok_group_max(const float* in, int64_t ngroups, int64_t k, int64_t C, float* out, void* stream) {
  P3_REQUIRE(ngroups >= 0 && k > 0 && C > 0, P3TOK_ERR_INVALID, "group_max: bad shape");
  if (ngroups == 0) return P3TOK_OK;
  P3_REQUIRE(in && out, P3TOK_ERR_INVALID, "group_max: null pointer");
  return group_max(in, ngroups, k, C, 0, out, as_stream(stream));
}

// cls rows of the token head: feats[b,0,:] = cls_token, pos[b,0,:] = cls_pos
__global__ void cls_rows_kernel(const float* __restrict__ cls_token, const float* __restrict__ cls_pos, int64_t B, int64_t G,
                                int E, float* __restrict__ feats, float* __restrict__ pos) {
  const int64_t total = B * E;
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = e / E;
    const int c = (int)(e - b * E);
    feats[b * (G + 1) * E + c] = cls_token[c];
    pos[b * (G + 1) * E + c] = cls_pos[c];
  }
}

extern "C" int p3tok_token_head_f32(const float* tokens, const float* centres, int64_t B, int64_t G, int64_t W, int64_t E,
                                    int64_t H, const float* proj_w, const float* proj_b, const float* pos_w1,
                                    const float* pos_b1, const float* pos_w2, const float* pos_b2, const float* cls_token,
                                    const float* cls_pos, float* hidden_ws, float* feats_out, float* pos_out, void* stream) {
  P3_REQUIRE(B >= 0 && G > 0 && W > 0 && E > 0 && H > 0 && G < (1ll << 30), P3TOK_ERR_INVALID, "token_head: bad shape");
  if (B == 0) return P3TOK_OK;
  P3_REQUIRE(tokens && centres && proj_w && pos_w1 && pos_w2 && cls_token && cls_pos && hidden_ws && feats_out && pos_out,
             P3TOK_ERR_INVALID, "token_head: null pointer");
  cudaStream_t s = as_stream(stream);
  const int64_t M = B * G;
  // x = proj(tokens) written behind each cloud's cls row (pix4point.py:245, 248)
  int rc = linear_f32(tokens, M, W, proj_w, E, proj_b, nullptr, 1, 0, feats_out, s, (int)G, 1);
  if (rc) return rc;
  // pos_embed = Linear(H->E)(GELU(Linear(3->H)(centres)))  (pix4point.py:214-218, 246)
  rc = linear_f32(centres, M, 3, pos_w1, H, pos_b1, nullptr, 1, 2, hidden_ws, s);
  if (rc) return rc;
  rc = linear_f32(hidden_ws, M, H, pos_w2, E, pos_b2, nullptr, 1, 0, pos_out, s, (int)G, 1);
  if (rc) return rc;
  cls_rows_kernel<<<grid_1d(B * E, 256), 256, 0, s>>>(cls_token, cls_pos, B, G, (int)E, feats_out, pos_out);
  P3_LAUNCH_CHECK("cls_rows_kernel");
  return P3TOK_OK;
}

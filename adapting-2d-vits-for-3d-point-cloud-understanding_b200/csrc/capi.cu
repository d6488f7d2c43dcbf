// capi.cu - error plumbing and the precision dispatch of the C ABI (include/p3tok.h).
#include <stdarg.h>
#include <atomic>
#include <string.h>

#include "embed.cuh"

namespace p3tok {

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what) {
  set_error("CUDA error %d (%s) at %s", (int)e, cudaGetErrorString(e), what);
  return P3TOK_ERR_CUDA;
}

static int check_embed_args(const p3tok_rows* R, const p3tok_mlp* m) {
  P3_REQUIRE(R && m, P3TOK_ERR_INVALID, "patch_embed: null descriptor");
  P3_REQUIRE(R->kind >= 0 && R->kind <= 2, P3TOK_ERR_INVALID, "patch_embed: rows.kind %d", R->kind);
  P3_REQUIRE(R->B >= 0 && R->G >= 0 && R->k > 0, P3TOK_ERR_INVALID, "patch_embed: bad B/G/k");
  P3_REQUIRE(m->n_pre >= 1 && m->n_pre <= 4, P3TOK_ERR_INVALID, "patch_embed: n_pre %d", m->n_pre);
  const int cin = R->kind == 0 ? 2 * R->C : (R->kind == 1 ? 3 + R->D : m->cin);
  P3_REQUIRE(cin == m->cin && cin > 0, P3TOK_ERR_INVALID, "patch_embed: rows give %d channels, mlp.cin=%d", cin, m->cin);
  P3_REQUIRE(m->mid_dim > 0 && m->out_dim > 0, P3TOK_ERR_INVALID, "patch_embed: bad widths");
  for (int i = 0; i < m->n_pre; ++i)
    P3_REQUIRE(m->pre_dim[i] > 0 && m->w_pre[i], P3TOK_ERR_INVALID, "patch_embed: layer %d missing", i);
  P3_REQUIRE(m->w_mid_g && m->w_mid_f && m->w_out, P3TOK_ERR_INVALID, "patch_embed: null weights");
  if (R->B * R->G > 0) {
    P3_REQUIRE(R->x && (R->kind == 2 || R->knn_idx), P3TOK_ERR_INVALID, "patch_embed: null rows pointer");
    P3_REQUIRE(R->kind != 0 || (R->ctr_idx && R->C >= 3), P3TOK_ERR_INVALID, "patch_embed: APF rows need ctr_idx, C>=3");
    P3_REQUIRE(R->kind != 1 || (R->feats && R->C == 3), P3TOK_ERR_INVALID, "patch_embed: P4P rows need feats, C==3");
    P3_REQUIRE(R->kind == 2 || R->idx_dtype == P3TOK_I64 || R->idx_dtype == P3TOK_I32, P3TOK_ERR_INVALID,
               "patch_embed: idx dtype");
  }
  return P3TOK_OK;
}

}  // namespace p3tok

using namespace p3tok;

extern "C" int p3tok_abi_version(void) { return P3TOK_ABI_VERSION; }
extern "C" const char* p3tok_last_error(void) { return g_err; }
extern "C" int64_t p3tok_kernel_launches(void) { return (int64_t)g_launches.load(std::memory_order_relaxed); }

extern "C" int64_t p3tok_patch_embed_workspace_bytes(const p3tok_mlp* mlp, int64_t ngroups, int64_t k, int precision) {
  if (!mlp || ngroups < 0 || k <= 0 || mlp->n_pre < 1 || mlp->n_pre > 4) return -1;
  if (precision == P3TOK_F32) return patch_embed_f32_workspace(mlp, ngroups, k);
  if (precision == P3TOK_BF16) return patch_embed_bf16_workspace(mlp, ngroups, k);
  if (precision == P3TOK_BF16X3) return patch_embed_x3_workspace(mlp, ngroups, k);
  return -1;
}

extern "C" int p3tok_patch_embed(const p3tok_rows* rows, const p3tok_mlp* mlp, int precision, int tokens_dtype,
                                 void* workspace, int64_t workspace_bytes, void* tokens, void* stream) {
  int rc = check_embed_args(rows, mlp);
  if (rc) return rc;
  if (rows->B * rows->G == 0) return P3TOK_OK;
  P3_REQUIRE(tokens && workspace, P3TOK_ERR_INVALID, "patch_embed: null output/workspace");
  P3_REQUIRE(tokens_dtype == P3TOK_F32 || tokens_dtype == P3TOK_BF16, P3TOK_ERR_INVALID, "patch_embed: tokens_dtype %d", tokens_dtype);
  P3_REQUIRE(tokens_dtype == P3TOK_F32 || precision == P3TOK_BF16, P3TOK_ERR_UNSUPPORTED,
             "patch_embed: bf16 tokens are emitted by the bf16 path only");
  if (precision == P3TOK_F32) {
    P3_REQUIRE(mlp->wdtype == P3TOK_F32, P3TOK_ERR_INVALID, "patch_embed(f32): weights must be f32");
    return patch_embed_f32(rows, mlp, workspace, workspace_bytes, (float*)tokens, as_stream(stream));
  }
  if (precision == P3TOK_BF16) {
    P3_REQUIRE(mlp->wdtype == P3TOK_BF16, P3TOK_ERR_INVALID, "patch_embed(bf16): weights must be bf16");
    return patch_embed_bf16(rows, mlp, workspace, workspace_bytes, tokens, tokens_dtype == P3TOK_BF16, as_stream(stream));
  }
  if (precision == P3TOK_BF16X3) {
    P3_REQUIRE(mlp->wdtype == P3TOK_BF16X3, P3TOK_ERR_INVALID, "patch_embed(bf16x3): weights must be the split [hi|hi|lo] form");
    return patch_embed_x3(rows, mlp, workspace, workspace_bytes, (float*)tokens, as_stream(stream));
  }
  set_error("patch_embed: unknown precision %d", precision);
  return P3TOK_ERR_INVALID;
}

// embed.cuh - internal interfaces shared by the patch-embedding paths.
#pragma once
#include "common.cuh"

namespace p3tok {

// rows of the first layer's input, built from the fused-gather description (p3tok_rows).
// X: (ngroups_chunk*k, cin) f32 for groups [g_begin, g_begin+g_count) in OUTPUT (permuted) order.
int build_rows_f32(const p3tok_rows* rows, int64_t g_begin, int64_t g_count, float* X, cudaStream_t s);

int64_t patch_embed_f32_workspace(const p3tok_mlp* mlp, int64_t ngroups, int64_t k);
int patch_embed_f32(const p3tok_rows* rows, const p3tok_mlp* mlp, void* ws, int64_t ws_bytes, float* tokens,
                    cudaStream_t s);

int64_t patch_embed_bf16_workspace(const p3tok_mlp* mlp, int64_t ngroups, int64_t k);
int patch_embed_bf16(const p3tok_rows* rows, const p3tok_mlp* mlp, void* ws, int64_t ws_bytes, float* tokens,
                     cudaStream_t s);

static inline int64_t align_up(int64_t v, int64_t a) { return (v + a - 1) / a * a; }

}  // namespace p3tok

// embed_tc.cu - bf16 tcgen05 patch embedding (placeholder until the tensor-core path lands).
#include "embed.cuh"
namespace p3tok {
int64_t patch_embed_bf16_workspace(const p3tok_mlp*, int64_t, int64_t) { return 256; }
int patch_embed_bf16(const p3tok_rows*, const p3tok_mlp*, void*, int64_t, float*, cudaStream_t) {
  set_error("patch_embed(bf16): not built yet");
  return P3TOK_ERR_UNSUPPORTED;
}
}  // namespace p3tok

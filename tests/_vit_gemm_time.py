"""Scratch: per-kernel timing of the ViT layer's building blocks (CUDA events, 20 iterations, warm L2).
python tests/_vit_gemm_time.py [M]     P3TOK_TC_TRACE=1 prints CTA 0's timeline of every GEMM launch (use with M small runs)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "adapting-2d-vits-for-3d-point-cloud-understanding_b200")]
import torch
from p3tok import ops

M = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
D, H, R, G, heads = 384, 1536, 64, 128, 12
dev = torch.device("cuda:0")
torch.manual_seed(0)
bf = lambda *s: (torch.randn(*s, device=dev) * 0.05).bfloat16()
x = torch.randn(M, D, device=dev)
a, o = bf(M, D), bf(M, D)
qkv, h = bf(M, 3 * D), bf(M, H + R)
w_qkv, w_proj, w_fc1d, w_fc2u = bf(3 * D, D), bf(D, D), bf(H + R, D), bf(D, H + R)
b3, b1, bh = torch.zeros(3 * D, device=dev), torch.zeros(D, device=dev), torch.zeros(H + R, device=dev)
trace = bool(os.environ.get("P3TOK_TC_TRACE"))
cases = [
    ("layernorm", lambda: ops.layernorm_bf16(x, None, None, 1e-5)),
    ("qkv  K=384 N=1152 bf16", lambda: ops.linear_bf16_ex(a, w_qkv, b3, 0, 0, None, 0.0, 1.0)),
    ("attention", lambda: ops.attention_bf16(qkv, M // G, G, heads)),
    ("proj K=384 N=384 +res", lambda: ops.linear_bf16_ex(o, w_proj, b1, 0, 0, x, 1.0, 1.0)),
    ("proj K=384 N=384 bf16 (no res)", lambda: ops.linear_bf16_ex(o, w_proj, b1, 0, 0, None, 0.0, 1.0)),
    ("fc1d K=384 N=1600 gelu|relu", lambda: ops.linear_bf16_ex(a, w_fc1d, bh, 3, H, None, 0.0, 1.0)),
    ("fc1d K=384 N=1600 no act", lambda: ops.linear_bf16_ex(a, w_fc1d, bh, 0, 0, None, 0.0, 1.0)),
    ("fc2u K=1600 N=384 +res", lambda: ops.linear_bf16_ex(h, w_fc2u, b1, 0, 0, x, 2.0, 1.0)),
    ("fc2u K=1600 N=384 bf16 (no res)", lambda: ops.linear_bf16_ex(h, w_fc2u, b1, 0, 0, None, 0.0, 1.0)),
]
for name, fn in cases:
    if trace:
        print("=====", name, file=sys.stderr)
        fn()
        continue
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        fn()
    e1.record()
    torch.cuda.synchronize()
    print(f"M={M} {name:36s} {e0.elapsed_time(e1) / 20 * 1e3:8.1f} us (includes the op's output allocation)")

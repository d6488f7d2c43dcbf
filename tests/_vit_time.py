"""Scratch timing of the ViT block stack at BASELINE config 2 (B=128, G=128, ViT-S): python tests/_vit_time.py [B G D heads]"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "adapting-2d-vits-for-3d-point-cloud-understanding_b200")]
import torch
from p3tok import synth, ops
from p3tok.apf_model import APFViTLayer, run_blocks

B, G, D, heads = (int(v) for v in (sys.argv[1:5] if len(sys.argv) >= 5 else (128, 128, 384, 12)))
depth = 12
dev = torch.device("cuda:0")
sd = synth.to_torch_state(synth.apf_vit_state(D, depth, 15, 5))
blocks = torch.nn.Sequential(*[APFViTLayer(D, heads) for _ in range(depth)]).eval().to(dev)
blocks.load_state_dict({k[len("blocks."):]: v for k, v in sd.items() if k.startswith("blocks.")})
norm = torch.nn.LayerNorm(D).eval().to(dev)
x = torch.randn(B, G, D, device=dev)
cache = {}
for _ in range(3):
    run_blocks(blocks, x, norm, cache)
torch.cuda.synchronize()
n0 = ops.kernel_launches()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
iters = 10
e0.record()
for _ in range(iters):
    run_blocks(blocks, x, norm, cache)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / iters
M = B * G
flops = depth * (2 * M * (D * 3 * D + D * D + 2 * D * 4 * D + 2 * D * 64) + 4 * B * heads * G * G * (D // heads))
print(f"vit stack B={B} G={G} D={D}: {ms:.3f} ms/forward, {B / ms * 1e3:.0f} clouds/s, {flops / ms / 1e9:.0f} TFLOP/s, "
      f"{(ops.kernel_launches() - n0) // iters} launches")

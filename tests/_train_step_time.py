"""Scratch timing of one APF training step on the kernels (AdaptPointFormer.train(), the reference's _freeze rule, cross-entropy,
backward, SGD step) at BASELINE config 2 shapes: python tests/_train_step_time.py [B]   (default 32, the reference's batch size)"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "adapting-2d-vits-for-3d-point-cloud-understanding_b200")]
import torch
from p3tok import ops, synth
from p3tok.apf_model import AdaptPointFormer

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
N, G, k, E = 2048, 128, 32, 384
dev = torch.device("cuda:0")
m = AdaptPointFormer(num_classes=15, embedding_dim=E, npoint=G, nsample=k, in_channels=3, precision="fp32").to(dev).train()
m._freeze()
opt = torch.optim.SGD([p for p in m.parameters() if p.requires_grad], lr=1e-3)
x = torch.from_numpy(synth.make_cloud("uniform", B, N, 3, 3)).to(dev)
st = torch.from_numpy(synth.start_indices(B, N, 3)).to(dev)
y = torch.arange(B, device=dev) % 15


def step():
    opt.zero_grad(set_to_none=True)
    loss = torch.nn.functional.cross_entropy(m(x, st), y)
    loss.backward()
    opt.step()
    return loss


for _ in range(2):
    step()
torch.cuda.synchronize()
n0 = ops.kernel_launches()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
iters = 5
t0 = time.perf_counter()
e0.record()
for _ in range(iters):
    loss = step()
e1.record()
torch.cuda.synchronize()
import os
print(f"[P3TOK_TRAIN_TC={os.environ.get('P3TOK_TRAIN_TC', '0')}] APF training step B={B} N={N} G={G} k={k} E={E} depth=12: {e0.elapsed_time(e1) / iters:.1f} ms/step (device), "
      f"{(time.perf_counter() - t0) / iters * 1e3:.1f} ms wall, {B / (e0.elapsed_time(e1) / iters) * 1e3:.0f} clouds/s, "
      f"{(ops.kernel_launches() - n0) // iters} p3tok launches, loss {float(loss):.4f}, peak memory {torch.cuda.max_memory_allocated() / 2**30:.2f} GiB")

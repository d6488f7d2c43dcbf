"""Scratch: kernel-time breakdown of one APF training step (torch.profiler, CUDA activities): python tests/_train_step_profile.py [B]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "adapting-2d-vits-for-3d-point-cloud-understanding_b200")]
import torch
from torch.profiler import ProfilerActivity, profile
from p3tok import synth
from p3tok.apf_model import AdaptPointFormer

B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
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
    torch.nn.functional.cross_entropy(m(x, st), y).backward()
    opt.step()


step(); step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    step()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=60))

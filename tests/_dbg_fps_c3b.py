import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "adapting-2d-vits-for-3d-point-cloud-understanding_b200"))
import numpy as np, torch
from oracle import oracle
from p3tok import ops, synth
dev = torch.device("cuda:0")
B, N = 2, 8192
p = synth.make_cloud("uniform", B, N, 3300 + 7, 3)
starts = [synth.start_indices(B, N, 33, 0), synth.start_indices(B, N // 4, 33, 1)]
f0 = oracle.fps(p, starts[0], N // 4)
c0 = oracle.gather_points(p, f0)[1:2].copy()          # cloud 1 alone: (1, 2048, 3)
s1 = starts[1][1:2].copy()
f1 = oracle.fps(c0, s1, 512)[0]
ct = torch.from_numpy(c0).to(dev)
g = ops.fps(ct, torch.from_numpy(s1).to(dev), 512).cpu().numpy()[0]
print("first mismatch", int(np.argmin(g == f1)), g[170:176], f1[170:176])
# numpy float32 running distance after 173 picks
sel = c0[0][f1[:173]]
md = np.full(2048, 1e10, np.float32)
for c in sel:
    d = c0[0] - c
    d = d * d
    md = np.minimum(md, (d[:, 0] + d[:, 1]) + d[:, 2])
print("numpy md[39], md[218] bits:", md[39].view(np.uint32), md[218].view(np.uint32), "argmax", md.argmax(), "n at max", int((md == md.max()).sum()))
print("points 39 / 218:", c0[0][39], c0[0][218])
# which centre gives the min for each
for i in (39, 218):
    d = c0[0][i] - sel
    d = d * d
    dd = (d[:, 0] + d[:, 1]) + d[:, 2]
    print(i, "nearest selected iteration", int(dd.argmin()), dd.min().view(np.uint32))
# the same on the GPU with separate torch kernels (no contraction possible)
x = ct[0]
mdt = torch.full((2048,), 1e10, device=dev)
for c in torch.from_numpy(sel).to(dev):
    d = x - c
    d = d * d
    mdt = torch.minimum(mdt, (d[:, 0] + d[:, 1]) + d[:, 2])
print("torch-cuda md bits:", mdt[39].view(torch.int32).item(), mdt[218].view(torch.int32).item(), "argmax", int(mdt.argmax()))
for G in (174, 175, 200):
    gg = ops.fps(ct, torch.from_numpy(s1).to(dev), G).cpu().numpy()[0]
    print("G", G, "idx[173]", gg[173])
# start directly from a state: does a tie between 39 and 218 resolve to the lowest index?  two-point-tie synthetic cloud
y = np.zeros((1, 2048, 3), np.float32)
y[0, :, 0] = np.arange(2048) * 1e-4
y[0, 39] = (5.0, 0, 0); y[0, 218] = (-5.0, 0, 0); y[0, 0] = (0, 0, 0)
gg = ops.fps(torch.from_numpy(y).to(dev), torch.zeros(1, dtype=torch.long, device=dev), 3).cpu().numpy()[0]
print("synthetic tie 39 vs 218 ->", gg, "oracle", oracle.fps(y, np.zeros(1, np.int64), 3)[0])

import sys, os
sys.path.insert(0, 'adapting-2d-vits-for-3d-point-cloud-understanding_b200')
import torch, numpy as np
from p3tok import ops, synth
dev = torch.device('cuda:0')
x = torch.from_numpy(synth.make_cloud("uniform", 2309, 2048, 1, 3)).to(dev)
st = torch.zeros(2309, dtype=torch.long, device=dev)
ops.fps(x[:4], st[:4], 8); torch.cuda.synchronize()
for B in (128, 148, 296, 592, 1184, 2309):
    for G in (128, 1024):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); ops.fps(x[:B], st[:B], G); e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        print(f"B={B:5d} G={G:5d}: {ms:8.3f} ms  -> {1e3*ms/G:7.3f} us/iteration (whole grid)")
from p3tok import functional as F
xc = torch.from_numpy(synth.make_cloud("clustered", 2309, 2048, 77, 3)).to(dev)
stc = torch.from_numpy(synth.start_indices(2309, 2048, 77)).to(dev)
for rep in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); out = F.fps(xc, 1024, stc); e1.record(); torch.cuda.synchronize()
    print(f"F.fps full dataset rep {rep}: {e0.elapsed_time(e1):.3f} ms")
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); out = F.fps(xc, 2048, stc); e1.record(); torch.cuda.synchronize()
print(f"F.fps full dataset G=N=2048: {e0.elapsed_time(e1):.3f} ms")

"""Debug aid (GPU): stage-1 FPS of the C3-shape parity case against the oracle, with and without the side-stream overlap."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "adapting-2d-vits-for-3d-point-cloud-understanding_b200"))
import numpy as np, torch
from oracle import oracle
from p3tok import ops, synth

dev = torch.device("cuda:0")
B, N, k = 2, 8192, 32
for kind in ("uniform", "clustered"):
    p = synth.make_cloud(kind, B, N, 3300 + len(kind), 3)
    starts = [synth.start_indices(B, N, 33, 0), synth.start_indices(B, N // 4, 33, 1)]
    f0 = oracle.fps(p, starts[0], N // 4)
    c0 = oracle.gather_points(p, f0)
    f1 = oracle.fps(c0, starts[1], N // 16)
    pt, c0t = torch.from_numpy(p).to(dev), torch.from_numpy(c0).to(dev)
    s0, s1 = (torch.from_numpy(s).to(dev) for s in starts)
    g0 = ops.fps(pt, s0, N // 4).cpu().numpy()
    print(kind, "stage0 fps equal:", np.array_equal(g0, f0))
    for rep in range(3):
        g1 = ops.fps(c0t, s1, N // 16).cpu().numpy()
        eq = g1 == f1
        print(kind, "stage1 fps equal:", eq.all(), "first mismatch per cloud:", [int(np.argmin(e)) if not e.all() else -1 for e in eq],
              "n mismatch:", (~eq).sum(1))
    for b in range(B):
        g1b = ops.fps(c0t[b:b + 1].contiguous(), s1[b:b + 1], N // 16).cpu().numpy()
        print(kind, f"cloud {b} alone equal:", np.array_equal(g1b[0], f1[b]))
    gi, ws = ops.fps_with_knn_prepare(c0t, s1, N // 16)
    torch.cuda.synchronize()
    print(kind, "stage1 fps with overlap equal:", np.array_equal(gi.cpu().numpy(), f1))
    # where the first mismatch is: distances of the two candidates
    g1 = ops.fps(c0t, s1, N // 16).cpu().numpy()
    for b in range(B):
        e = g1[b] == f1[b]
        if e.all():
            continue
        i = int(np.argmin(e))
        sel = c0[b][f1[b][:i]]
        d = ((c0[b][:, None, :] - sel[None]) ** 2)
        dist = ((d[..., 0] + d[..., 1]) + d[..., 2]).min(1)
        print(kind, b, "iter", i, "oracle pick", f1[b][i], dist[f1[b][i]], "gpu pick", g1[b][i], dist[g1[b][i]], "max", dist.max(), int(dist.argmax()))

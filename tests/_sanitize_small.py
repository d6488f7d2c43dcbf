"""Small end-to-end run of every kernel for compute-sanitizer (memcheck / racecheck); not collected by pytest."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "adapting-2d-vits-for-3d-point-cloud-understanding_b200")); sys.path.insert(0, ROOT)
import numpy as np, torch
from p3tok import synth, ops, _lib
from p3tok.modules import PointNet, P3Embed, Group
dev = torch.device("cuda:0")
def t(a): return torch.from_numpy(np.ascontiguousarray(a)).to(dev)
for (B, N, G, k, C) in ((2, 300, 20, 8, 3), (1, 1100, 40, 32, 4), (1, 9000, 16, 64, 3)):
    x = t(synth.make_cloud("clustered", B, N, 3, C)); st = t(synth.start_indices(B, N, 3))
    idx = ops.fps(x, st, G); ctr = ops.gather_points(x, idx)
    for mode in (0, 1):
        ops.knn(x, ctr[..., :3].contiguous(), k, mode, mode == 1, True)
    Group(G, k)(x, x[:, :, :3], st)
for prec in ("fp32", "bf16"):
    net = PointNet(64, 16, 32, 6, precision=prec).eval().to(dev)
    net.encoder.load_state_dict(synth.to_torch_state(synth.apf_encoder_state(64, 6, 1)))
    x = t(synth.make_cloud("uniform", 3, 512, 1, 3)); st = t(synth.start_indices(3, 512, 1))
    out = net(x, st)
    m = P3Embed(sample_ratio=1 / 16, k=16, embed_dim=128, precision=prec).eval().to(dev)
    m.load_state_dict(synth.to_torch_state(synth.p3embed_state(3, 1 / 16, 4, 4, 128, 2)))
    ps, fs = m(x, x.transpose(1, 2).contiguous(), [st, t(synth.start_indices(3, 128, 1, 1))])
pn = t(synth.make_points_nd(2, 5000, 6, 4))                      # general-D FPS (both loop bodies) and the squared-distance matrix
ops.fps_nd(pn, t(synth.start_indices(2, 5000, 4)), 12); ops.fps_nd(pn[:, :300].contiguous(), t(synth.start_indices(2, 300, 4)), 12)
ops.square_distance(x[:, :37].contiguous(), x); ops.square_distance(x[:, :5, :3].contiguous(), x[:, :301].contiguous())
torch.cuda.synchronize()
print("sanitize-run ok", float(out.abs().sum()), float(fs[-1].abs().sum()))

"""Scratch driver (not a test): APF tokens at C2 / C4-like shapes with the first layer inside the pair kernel
(P3TOK_L1_FUSED=1, default) vs the separate first-layer kernel (=0) - the two must agree bit for bit.
usage: python tests/_l1_ab.py dump <tag>   |   python tests/_l1_ab.py cmp <tagA> <tagB>"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "adapting-2d-vits-for-3d-point-cloud-understanding_b200"))
OUT = os.path.join(ROOT, "gpurun_out")

CASES = [(128, 2048, 3, 128, 32, 384), (5, 1024, 4, 37, 32, 384), (2, 16384, 3, 512, 64, 384), (3, 512, 3, 33, 32, 128)]

if sys.argv[1] == "dump":
    import torch
    from p3tok import synth
    from p3tok.modules import PointNet
    dev = torch.device("cuda:0")
    for ci, (B, N, C, G, k, E) in enumerate(CASES):
        x = synth.make_cloud("clustered", B, N, 500 + ci, C)
        start = synth.start_indices(B, N, 500 + ci)
        sd = synth.apf_encoder_state(E, 2 * C, 500 + ci)
        net = PointNet(E, G, k, 2 * C, precision="bf16").eval().to(dev)
        net.encoder.load_state_dict(synth.to_torch_state(sd))
        tok = net(torch.from_numpy(x).to(dev), torch.from_numpy(start).to(dev))
        torch.cuda.synchronize()
        np.save(os.path.join(OUT, f"l1ab_{sys.argv[2]}_{ci}.npy"), tok.float().cpu().numpy())
        print("dumped", ci, tok.shape, float(tok.abs().max()))
else:
    ok = True
    for ci in range(len(CASES)):
        a = np.load(os.path.join(OUT, f"l1ab_{sys.argv[2]}_{ci}.npy"))
        b = np.load(os.path.join(OUT, f"l1ab_{sys.argv[3]}_{ci}.npy"))
        same = np.array_equal(a, b)
        print("case", ci, CASES[ci], "bit-identical" if same else f"DIFFER max {np.abs(a - b).max():.3e} of {np.abs(a).max():.3e}")
        ok &= same
    sys.exit(0 if ok else 1)

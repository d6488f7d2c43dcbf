"""Debug helper (not a test): one APF PointNet forward at the c2 shape; with P3TOK_TC_TRACE=1 the library prints the
per-tile timeline of every tensor-core GEMM."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "adapting-2d-vits-for-3d-point-cloud-understanding_b200"))
import torch
from p3tok import synth
from p3tok.modules import PointNet
B, N, G, k, E = 128, 2048, 128, 32, 384
net = PointNet(E, G, k, 6, precision="bf16").eval().cuda()
net.encoder.load_state_dict(synth.to_torch_state(synth.apf_encoder_state(E, 6, 0)))
x = torch.from_numpy(synth.make_cloud("uniform", B, N, 1, 3)).cuda()
st = torch.from_numpy(synth.start_indices(B, N, 1)).cuda()
os.environ.pop("P3TOK_TC_TRACE_OFF", None)
tok = net(x, st)
torch.cuda.synchronize()

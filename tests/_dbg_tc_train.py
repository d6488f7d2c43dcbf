"""Scratch: per-GEMM error of the bf16x3 training products against float64, on the real operands of an Encoder training step."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "adapting-2d-vits-for-3d-point-cloud-understanding_b200"), os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "golden")]
import numpy as np, torch
from oracle import oracle
from p3tok import synth, train
from p3tok.modules import Encoder

dev = torch.device("cuda:0")
B, N, C, G, k, E = 3, 512, 3, 43, 16, 64
xs = synth.make_cloud("uniform", B, N, 77, C)
neigh = oracle.group_apf(xs, synth.start_indices(B, N, 77), G, k)["neigh"].astype(np.float32)
sd = synth.apf_encoder_state(E, 2 * C, 77)
gt = (synth.uniform01(77, B * G * E, 31).reshape(B, G, E) - 0.5).astype(np.float32)
enc = Encoder(E, 2 * C).to(dev).train()
enc.load_state_dict(synth.to_torch_state(sd))
real = train._linear_x3
def logged(a, w, b, g, rpg):
    out = real(a, w, b, g, rpg)
    ref = a.double() @ w.double().t()
    if b is not None: ref = ref + b.double()
    if g is not None: ref = ref + g.double().repeat_interleave(rpg, 0)
    err = (out.double() - ref)
    nz = float((a != 0).float().mean())
    print(f"x3 {tuple(a.shape)} x {tuple(w.shape)}: max|err|/max|ref| {float(err.abs().max() / ref.abs().max()):.2e}  fro {float(err.norm() / ref.norm()):.2e}  "
          f"|a| max {float(a.abs().max()):.2e} nonzero share {nz:.3f} max|ref| {float(ref.abs().max()):.2e}")
    return out
train._linear_x3 = logged
res = {}
for lvl in (0, 1):
    for p_ in enc.parameters(): p_.grad = None
    x = torch.from_numpy(neigh).to(dev).requires_grad_(True)
    train.set_tensor_core_gemms(lvl)
    tok = enc(x)
    (tok * torch.from_numpy(gt).to(dev)).sum().backward()
    res[lvl] = (tok.detach().double(), x.grad.double(), {n: p_.grad.double().clone() for n, p_ in enc.named_parameters()})
rel = lambda a, b: float((a - b).norm() / b.norm())
print("tokens tc vs f32 fro", rel(res[1][0], res[0][0]), " grad input fro", rel(res[1][1], res[0][1]))
for n in res[0][2]:
    print(n, rel(res[1][2][n], res[0][2][n]) if float(res[0][2][n].norm()) > 0 else "zero")
# sensitivity of the function itself: fp32 SGEMM path, first-layer weights perturbed by 1e-5 relative noise
train.set_tensor_core_gemms(0)
with torch.no_grad():
    w = enc.first_conv[3].weight
    w.mul_(1.0 + 1e-5 * torch.randn_like(w))
for p_ in enc.parameters(): p_.grad = None
x = torch.from_numpy(neigh).to(dev).requires_grad_(True)
tok = enc(x)
(tok * torch.from_numpy(gt).to(dev)).sum().backward()
print("fp32 path, weights perturbed by 1e-5: tokens fro", rel(tok.detach().double(), res[0][0]), " grad input fro", rel(x.grad.double(), res[0][1]),
      " grad second_conv.3.weight fro", rel(enc.second_conv[3].weight.grad.double(), res[0][2]["second_conv.3.weight"]))

"""Shared helpers for the tests (test infrastructure)."""
import numpy as np
import torch


def dev():
    return torch.device("cuda:0")


def to_dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(dev())


def rel_err(a, ref):
    a = np.asarray(a, np.float64)
    ref = np.asarray(ref, np.float64)
    return float(np.abs(a - ref).max() / max(np.abs(ref).max(), 1e-30))


PARITY_LOG = []   # one record per assert_tokens_close call; tests/conftest.py writes it to gpurun_out/parity_report.json


def assert_tokens_close(a, ref, rtol, what=""):
    """The north-star tolerance (rtol 1e-4 for the fp32 path, 1e-2 for the bf16 path), written out:
         elementwise   |a - ref| <= rtol * |ref| + rtol * max|ref|     (allclose with a scale-aware atol:
                       ReLU/max-pooled tokens that cancel to ~0 cannot meet a pure relative bound in any
                       reduced precision)
         aggregate     ||a - ref||_F <= 0.5 * rtol * ||ref||_F
    How far that is from a PURE relative bound is measured, not argued: every call records the share of elements with
    |a - ref| <= rtol * |ref| (all elements, and those with |ref| >= 1 % of max|ref|, i.e. not cancelled to ~0) in
    PARITY_LOG, printed with `pytest -s` and written to gpurun_out/parity_report.json at session end.
    """
    a = np.asarray(a, np.float64)
    ref = np.asarray(ref, np.float64)
    assert a.shape == ref.shape, (a.shape, ref.shape)
    assert np.isfinite(a).all(), f"{what}: non-finite tokens"
    diff = np.abs(a - ref)
    amax = max(float(np.abs(ref).max()), 1e-30)
    pure = diff <= rtol * np.abs(ref)
    big = np.abs(ref) >= 1e-2 * amax
    fro = float(np.linalg.norm(a - ref) / max(np.linalg.norm(ref), 1e-30))
    rec = dict(what=what, rtol=rtol, n=int(ref.size), max_err_over_max=float(diff.max() / amax), fro=fro,
               share_pure_rtol=float(pure.mean()), share_pure_rtol_nonsmall=float(pure[big].mean()) if big.any() else 1.0)
    PARITY_LOG.append(rec)
    print(f"[parity] {what}: rtol {rtol:g}  max|err|/max|ref| {rec['max_err_over_max']:.2e}  fro {fro:.2e}  "
          f"pure-rtol share {rec['share_pure_rtol']:.4f} (|ref| >= 1% of max: {rec['share_pure_rtol_nonsmall']:.4f})")
    bound = rtol * np.abs(ref) + rtol * amax
    bad = diff > bound
    assert not bad.any(), (f"{what}: {int(bad.sum())}/{bad.size} tokens outside rtol={rtol}; "
                           f"worst ratio {float((diff / bound).max()):.2f}")
    assert fro <= 0.5 * rtol, f"{what}: relative Frobenius error {fro:.2e} > {0.5 * rtol:.1e}"
    return rec


def bf16_emulation(mlp, rows, k):
    """The bf16 path's arithmetic spelled in torch (any device): bf16 weights and stored activations,
    fp32 accumulation, max taken from the fp32 accumulators.  Test-side model of embed_tc.cu."""
    import torch

    def q(t):
        return t.bfloat16().float()

    h = rows.float()
    n = len(mlp.w_pre)
    for i, (w, b, r) in enumerate(zip(mlp.w_pre, mlp.b_pre, mlp.pre_relu)):
        if i > 0 or mlp.cin > 16:
            h = q(h)
        h = (h.double() @ q(w.to(h.device)).double().T + b.to(h.device).double()).float()
        if r:
            h = torch.relu(h)
    ng = h.shape[0] // k
    g = h.view(ng, k, -1).max(1)[0]
    dev_ = h.device
    gb = (q(g).double() @ q(mlp.w_mid_g.to(dev_)).double().T + mlp.b_mid.to(dev_).double()).float()
    h2 = torch.relu((q(h).double() @ q(mlp.w_mid_f.to(dev_)).double().T).float() + gb.repeat_interleave(k, 0))
    o = (q(h2).double() @ q(mlp.w_out.to(dev_)).double().T + mlp.b_out.to(dev_).double()).float()
    o = o.view(ng, k, -1).max(1)[0]
    return torch.relu(o) if mlp.out_relu else o


def folded_forward(mlp, rows, k):
    """float64 evaluation of a folded p3tok.fold.PatchMLP - checks the folding algebra."""
    h = rows.double()
    for w, b, r in zip(mlp.w_pre, mlp.b_pre, mlp.pre_relu):
        h = h @ w.double().T + b.double()
        if r:
            h = torch.relu(h)
    ng = h.shape[0] // k
    g = h.view(ng, k, -1).max(dim=1)[0]
    gb = g @ mlp.w_mid_g.double().T + mlp.b_mid.double()
    h = torch.relu(h @ mlp.w_mid_f.double().T + gb.repeat_interleave(k, dim=0))
    o = (h @ mlp.w_out.double().T + mlp.b_out.double()).view(ng, k, -1).max(dim=1)[0]
    return torch.relu(o) if mlp.out_relu else o


def vit_bf16_emulation(folded, tokens, heads, bottleneck, final_w, final_b):
    """The ViT block stack's arithmetic (csrc/vit.cu) spelled in torch on any device, from the FOLDED layers
    (p3tok.apf_model.fold_vit_layer: 8 tensors per layer): bf16 weights and stored activations (normalised rows, qkv,
    softmax probabilities, attention output, [gelu | relu] hidden matrix), fp32 accumulation, fp32 residual stream.
    Returns (x (B,G,D), pooled (B,D)) float32."""
    import torch
    import torch.nn.functional as F

    def q(t):
        return t.bfloat16().float()

    mm = lambda a, w: (a.double() @ w.double().T).float()
    x = tokens.float()
    B, G, D = x.shape
    hd = D // heads
    for qkv_w, qkv_b, proj_w, proj_b, fc1d_w, fc1d_b, fc2u_w, fc2u_b in folded:
        a = q(F.layer_norm(x, (D,), None, None, 1e-5))
        qkv = q(mm(a, qkv_w) + qkv_b).reshape(B, G, 3, heads, hd).permute(2, 0, 3, 1, 4)
        s = (qkv[0].double() @ qkv[1].double().transpose(-2, -1)).float() * hd ** -0.5
        e = torch.exp(s - s.max(-1, keepdim=True)[0])
        o = (q(e).double() @ qkv[2].double()).float() / e.sum(-1, keepdim=True)     # P rounded to bf16, row sum in fp32
        o = q(o.transpose(1, 2).reshape(B, G, D))
        x = x + (mm(o, proj_w) + proj_b)
        a = q(F.layer_norm(x, (D,), None, None, 1e-5))
        h = mm(a, fc1d_w) + fc1d_b
        H = h.shape[-1] - bottleneck
        h = q(torch.cat([F.gelu(h[..., :H]), torch.relu(h[..., H:])], -1))
        x = 2.0 * x + (mm(h, fc2u_w) + fc2u_b)
    pooled = F.layer_norm(x, (D,), final_w.float(), final_b.float(), 1e-5).max(1)[0]
    return x, pooled

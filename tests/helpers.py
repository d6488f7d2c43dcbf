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


def assert_tokens_close(a, ref, rtol, what=""):
    """|a - ref| <= rtol * (|ref| + mean|ref|): the north-star tolerance (rtol 1e-4 fp32 / 1e-2 bf16)
    with an absolute floor of rtol x the mean token magnitude for entries that cancel to ~0."""
    a = np.asarray(a, np.float64)
    ref = np.asarray(ref, np.float64)
    assert a.shape == ref.shape, (a.shape, ref.shape)
    bound = rtol * (np.abs(ref) + np.abs(ref).mean())
    bad = np.abs(a - ref) > bound
    assert not bad.any(), (f"{what}: {int(bad.sum())}/{bad.size} tokens outside rtol={rtol}; "
                           f"worst ratio {float((np.abs(a - ref) / bound).max()):.2f}")


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

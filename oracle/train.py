"""Training-mode restatement of the APF Encoder (SURVEY.md 8f "next" #4) - TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Groundwork for the next row: float64 numpy forward with batch-statistics BatchNorm (reference src/models/apf.py:129-143 in
train mode: nn.BatchNorm1d, eps 1e-5, momentum 0.1, biased variance for the normalisation, unbiased for the running
estimate) and the backward of the whole mini-PointNet through both max-pools and the concat (apf.py:145-169), i.e. what
autograd computes for `Encoder.forward`.  Pinned against the reference module's own autograd by tests/golden/make_golden.py
(tests/golden/apf_train.npz) and re-checked against those fixtures by tests/test_oracle_vs_golden.py.  The kernels held to
it: csrc/train.cu + p3tok/train.py (tokenizer), csrc/train_vit.cu + p3tok/train_vit.py (block stack, encoder_norm, head).
"""
from __future__ import annotations

from typing import Dict, Tuple

import numpy as np

BN_EPS = 1e-5
BN_MOMENTUM = 0.1


def _w(sd, name):
    w = np.asarray(sd[name + ".weight"], np.float64)
    return w.reshape(w.shape[0], -1)


def _bn_fwd(z, gamma, beta):
    mu = z.mean(0)
    var = z.var(0)                       # biased: what the normalisation uses
    xhat = (z - mu) / np.sqrt(var + BN_EPS)
    return gamma * xhat + beta, (xhat, var, mu)


def _bn_bwd(dy, gamma, cache):
    xhat, var, _ = cache
    dgamma = (dy * xhat).sum(0)
    dbeta = dy.sum(0)
    dz = gamma / np.sqrt(var + BN_EPS) * (dy - dy.mean(0) - xhat * (dy * xhat).mean(0))
    return dz, dgamma, dbeta


def apf_encoder_train(sd: Dict[str, np.ndarray], neigh: np.ndarray, grad_tokens: np.ndarray):
    """neigh (B,G,k,cin), grad_tokens (B,G,E) = dL/dtokens.
    Returns (tokens (B,G,E), grads: dict name -> array incl. "input", running: dict of updated BN running statistics)."""
    B, G, k, cin = neigh.shape
    f = lambda n: np.asarray(sd[n], np.float64)
    X = neigh.reshape(B * G * k, cin).astype(np.float64)
    R = X.shape[0]
    W1, W2, W3, W4, W5 = (_w(sd, n) for n in ("first_conv.0", "first_conv.3", "first_conv.6", "second_conv.0", "second_conv.3"))
    b1, b2, b3, b4, b5 = (f(n + ".bias") for n in ("first_conv.0", "first_conv.3", "first_conv.6", "second_conv.0", "second_conv.3"))
    g1, be1, g2, be2, g4, be4 = (f(n) for n in ("first_conv.1.weight", "first_conv.1.bias", "first_conv.4.weight", "first_conv.4.bias",
                                                "second_conv.1.weight", "second_conv.1.bias"))
    # ---- forward (apf.py:145-169)
    z1 = X @ W1.T + b1
    y1, c1 = _bn_fwd(z1, g1, be1)
    h1 = np.maximum(y1, 0)
    z2 = h1 @ W2.T + b2
    y2, c2 = _bn_fwd(z2, g2, be2)
    h2 = np.maximum(y2, 0)
    ft = h2 @ W3.T + b3                                  # (R, E)
    E = ft.shape[1]
    ftg = ft.reshape(B * G, k, E)
    arg_g = ftg.argmax(1)                                # first maximum, like torch.max(dim)
    gl = np.take_along_axis(ftg, arg_g[:, None, :], 1)[:, 0]
    cat = np.concatenate([np.repeat(gl, k, 0), ft], 1)   # (R, 2E): [global | per-point]
    z4 = cat @ W4.T + b4
    y4, c4 = _bn_fwd(z4, g4, be4)
    h4 = np.maximum(y4, 0)
    o = h4 @ W5.T + b5
    og = o.reshape(B * G, k, E)
    arg_o = og.argmax(1)
    tokens = np.take_along_axis(og, arg_o[:, None, :], 1)[:, 0]
    running = {}
    for name, (_, var, mu) in (("first_conv.1", c1), ("first_conv.4", c2), ("second_conv.1", c4)):
        running[name + ".running_mean"] = (1 - BN_MOMENTUM) * f(name + ".running_mean") + BN_MOMENTUM * mu
        running[name + ".running_var"] = (1 - BN_MOMENTUM) * f(name + ".running_var") + BN_MOMENTUM * var * R / (R - 1)
    # ---- backward
    gt = grad_tokens.reshape(B * G, E).astype(np.float64)
    do = np.zeros_like(og)
    np.put_along_axis(do, arg_o[:, None, :], gt[:, None, :], 1)
    do = do.reshape(R, E)
    grads = {"second_conv.3.weight": do.T @ h4, "second_conv.3.bias": do.sum(0)}
    dy4 = (do @ W5) * (y4 > 0)
    dz4, grads["second_conv.1.weight"], grads["second_conv.1.bias"] = _bn_bwd(dy4, g4, c4)
    grads["second_conv.0.weight"] = dz4.T @ cat
    grads["second_conv.0.bias"] = dz4.sum(0)
    dcat = dz4 @ W4
    dgl = dcat[:, :E].reshape(B * G, k, E).sum(1)        # the expanded global feature collects its k copies
    dftg = dcat[:, E:].reshape(B * G, k, E).copy()
    add = np.zeros_like(dftg)
    np.put_along_axis(add, arg_g[:, None, :], dgl[:, None, :], 1)
    dft = (dftg + add).reshape(R, E)
    grads["first_conv.6.weight"] = dft.T @ h2
    grads["first_conv.6.bias"] = dft.sum(0)
    dy2 = (dft @ W3) * (y2 > 0)
    dz2, grads["first_conv.4.weight"], grads["first_conv.4.bias"] = _bn_bwd(dy2, g2, c2)
    grads["first_conv.3.weight"] = dz2.T @ h1
    grads["first_conv.3.bias"] = dz2.sum(0)
    dy1 = (dz2 @ W2) * (y1 > 0)
    dz1, grads["first_conv.1.weight"], grads["first_conv.1.bias"] = _bn_bwd(dy1, g1, c1)
    grads["first_conv.0.weight"] = dz1.T @ X
    grads["first_conv.0.bias"] = dz1.sum(0)
    grads["input"] = (dz1 @ W1).reshape(B, G, k, cin)
    return tokens.reshape(B, G, E), grads, running


def p3embed_stage_train(sd: Dict[str, np.ndarray], s: int, rows: np.ndarray, grad_out: np.ndarray):
    """One P3Embed stage in train mode on its gathered rows (reference src/models/pix4point.py:179-188 with the BatchNorm2d
    layers of 135-156 on batch statistics).  rows (B,G,k,Cin) = [grouped xyz || grouped feats], grad_out (B,G,W) = dL/d(stage
    output, channel-last).  Returns (out (B,G,W), grads: dict name -> array incl. "rows", running statistics).
    conv1 = Conv2d(Cin,W, no bias) -> Conv2d(W,W, bias) -> BN -> ReLU;  conv2 = Conv2d(2W,2W, no bias) -> BN -> ReLU ->
    Conv2d(2W,W, no bias) -> BN -> ReLU;  pools are max over k."""
    B, G, k, cin = rows.shape
    f = lambda n: np.asarray(sd[n], np.float64)
    pre = f"convs.{s}."
    Wa, Wb = _w(sd, pre + "0.0"), _w(sd, pre + "0.1")
    bb = f(pre + "0.1.bias")
    Wc, Wd = _w(sd, pre + "1.0"), _w(sd, pre + "1.3")
    g1, be1 = f(pre + "0.2.weight"), f(pre + "0.2.bias")
    g2, be2 = f(pre + "1.1.weight"), f(pre + "1.1.bias")
    g3, be3 = f(pre + "1.4.weight"), f(pre + "1.4.bias")
    X = rows.reshape(B * G * k, cin).astype(np.float64)
    R = X.shape[0]
    a = X @ Wa.T                                         # no bias, no activation (pix4point.py:139)
    z1 = a @ Wb.T + bb
    y1, c1 = _bn_fwd(z1, g1, be1)
    f1 = np.maximum(y1, 0)
    W = f1.shape[1]
    f1g = f1.reshape(B * G, k, W)
    arg_g = f1g.argmax(1)
    gl = np.take_along_axis(f1g, arg_g[:, None, :], 1)[:, 0]
    cat = np.concatenate([np.repeat(gl, k, 0), f1], 1)
    z2 = cat @ Wc.T
    y2, c2 = _bn_fwd(z2, g2, be2)
    h2 = np.maximum(y2, 0)
    z3 = h2 @ Wd.T
    y3, c3 = _bn_fwd(z3, g3, be3)
    h3 = np.maximum(y3, 0)
    h3g = h3.reshape(B * G, k, W)
    arg_o = h3g.argmax(1)
    out = np.take_along_axis(h3g, arg_o[:, None, :], 1)[:, 0]
    running = {}
    for name, (_, var, mu) in ((pre + "0.2", c1), (pre + "1.1", c2), (pre + "1.4", c3)):
        running[name + ".running_mean"] = (1 - BN_MOMENTUM) * f(name + ".running_mean") + BN_MOMENTUM * mu
        running[name + ".running_var"] = (1 - BN_MOMENTUM) * f(name + ".running_var") + BN_MOMENTUM * var * R / (R - 1)
    # ---- backward
    go = grad_out.reshape(B * G, W).astype(np.float64)
    dh3 = np.zeros_like(h3g)
    np.put_along_axis(dh3, arg_o[:, None, :], go[:, None, :], 1)
    dy3 = dh3.reshape(R, W) * (y3 > 0)
    grads = {}
    dz3, grads[pre + "1.4.weight"], grads[pre + "1.4.bias"] = _bn_bwd(dy3, g3, c3)
    grads[pre + "1.3.weight"] = dz3.T @ h2
    dy2 = (dz3 @ Wd) * (y2 > 0)
    dz2, grads[pre + "1.1.weight"], grads[pre + "1.1.bias"] = _bn_bwd(dy2, g2, c2)
    grads[pre + "1.0.weight"] = dz2.T @ cat
    dcat = dz2 @ Wc
    dgl = dcat[:, :W].reshape(B * G, k, W).sum(1)
    df1 = dcat[:, W:].reshape(B * G, k, W).copy()
    add = np.zeros_like(df1)
    np.put_along_axis(add, arg_g[:, None, :], dgl[:, None, :], 1)
    dy1 = (df1 + add).reshape(R, W) * (y1 > 0)
    dz1, grads[pre + "0.2.weight"], grads[pre + "0.2.bias"] = _bn_bwd(dy1, g1, c1)
    grads[pre + "0.1.weight"] = dz1.T @ a
    grads[pre + "0.1.bias"] = dz1.sum(0)
    da = dz1 @ Wb
    grads[pre + "0.0.weight"] = da.T @ X
    grads["rows"] = (da @ Wa).reshape(B, G, k, cin)
    return out.reshape(B, G, W), grads, running


def scatter_rows_grad(drows: np.ndarray, knn_idx: np.ndarray, N: int) -> Tuple[np.ndarray, np.ndarray]:
    """Gradient of the gather (pix4point.py:92-102): drows (B,G,k,3+D) -> (d points (B,N,3), d feats (B,N,D)), each point
    collecting the gradients of every neighbourhood it appears in."""
    B = drows.shape[0]
    D = drows.shape[-1] - 3
    dp, df = np.zeros((B, N, 3)), np.zeros((B, N, D))
    for b in range(B):
        idx = np.asarray(knn_idx[b]).reshape(-1)
        flat = drows[b].reshape(-1, 3 + D)
        np.add.at(dp[b], idx, flat[:, :3])
        np.add.at(df[b], idx, flat[:, 3:])
    return dp, df


# ------------------------------------------------------------------------------------------------------------------
# Backward THROUGH the (frozen) ViT block stack: what the tokenizer's Encoder needs from the layers above it when the
# reference trains (apf.py:335-346 keeps the blocks frozen but the gradient still flows through them).  dX only, plus the
# trainable encoder_norm.  Dropout / DropPath are taken at p = 0 (they are random in the reference's train mode).

def _erf(x):
    from scipy.special import erf
    return erf(x)


def _ln_fwd(x, w, b, eps=1e-5):
    mu = x.mean(-1, keepdims=True)
    var = ((x - mu) ** 2).mean(-1, keepdims=True)
    rstd = 1.0 / np.sqrt(var + eps)
    xhat = (x - mu) * rstd
    return xhat * w + b, (xhat, rstd)


def _ln_bwd(dy, w, cache):
    xhat, rstd = cache
    dxhat = dy * w
    dx = rstd * (dxhat - dxhat.mean(-1, keepdims=True) - xhat * (dxhat * xhat).mean(-1, keepdims=True))
    return dx, (dy * xhat).reshape(-1, xhat.shape[-1]).sum(0), dy.reshape(-1, xhat.shape[-1]).sum(0)


def apf_vit_layer_fwd_bwd(sd: Dict[str, np.ndarray], p: str, x: np.ndarray, heads: int, masks=None, grads: Dict[str, np.ndarray] = None):
    """Forward of one APFViTLayer (apf_utils.py:268-293) returning the output and a closure dY -> dX.
    masks = (drop_path_attn (B,), adapter dropout (B*G, R), drop_path_mlp (B,)) keep masks already scaled by 1 / keep
    probability (None entries = no dropout: the reference's eval mode or rate 0).  When `grads` is a dict the closure also
    fills it with the gradient of every parameter of the layer (state_dict names)."""
    f = lambda k: np.asarray(sd[p + k], np.float64)
    B, G, D = x.shape
    hd = D // heads
    dp1, dmask, dp2 = masks if masks is not None else (None, None, None)
    bc = lambda m: 1.0 if m is None else np.asarray(m, np.float64).reshape(B, 1, 1)
    dm3 = 1.0 if dmask is None else np.asarray(dmask, np.float64).reshape(B, G, -1)
    a, c_n1 = _ln_fwd(x, f("norm1.weight"), f("norm1.bias"))
    Wq, bq, Wp, bp = f("attention.qkv.weight"), f("attention.qkv.bias"), f("attention.proj.weight"), f("attention.proj.bias")
    qkv = (a @ Wq.T + bq).reshape(B, G, 3, heads, hd).transpose(2, 0, 3, 1, 4)
    q, k, v = qkv[0], qkv[1], qkv[2]
    s = (q @ k.transpose(0, 1, 3, 2)) * hd ** -0.5
    e = np.exp(s - s.max(-1, keepdims=True))
    P = e / e.sum(-1, keepdims=True)
    o = (P @ v).transpose(0, 2, 1, 3).reshape(B, G, D)
    x1 = x + (o @ Wp.T + bp) * bc(dp1)
    an, c_an = _ln_fwd(x1, f("adapter.adapter_norm.weight"), f("adapter.adapter_norm.bias"))
    Wd, bd, Wu, bu = f("adapter.down_proj.weight"), f("adapter.down_proj.bias"), f("adapter.up_proj.weight"), f("adapter.up_proj.bias")
    sc = float(np.asarray(sd[p + "adapter.scale"]).reshape(-1)[0])
    zd = an @ Wd.T + bd
    dn = np.maximum(zd, 0) * dm3
    up = dn @ Wu.T + bu
    n2, c_n2 = _ln_fwd(x1, f("norm2.weight"), f("norm2.bias"))
    W1, b1, W2, b2 = f("mlp.fc1.weight"), f("mlp.fc1.bias"), f("mlp.fc2.weight"), f("mlp.fc2.bias")
    z1 = n2 @ W1.T + b1
    cdf = 0.5 * (1.0 + _erf(z1 / np.sqrt(2.0)))
    h = z1 * cdf
    y = (h @ W2.T + b2) * bc(dp2) + (up * sc + x1) + x1

    def backward(dy):
        r2 = lambda t: t.reshape(-1, t.shape[-1])
        put = (lambda name, val: grads.__setitem__(p + name, val)) if grads is not None else (lambda name, val: None)
        dx1 = 2.0 * dy                                                    # adapter's "+ x" and the layer's
        dmm = dy * bc(dp2)
        put("mlp.fc2.weight", r2(dmm).T @ r2(h)); put("mlp.fc2.bias", r2(dmm).sum(0))
        dh = dmm @ W2
        dz1 = dh * (cdf + z1 * np.exp(-0.5 * z1 * z1) / np.sqrt(2.0 * np.pi))
        put("mlp.fc1.weight", r2(dz1).T @ r2(n2)); put("mlp.fc1.bias", r2(dz1).sum(0))
        d, gw, gb = _ln_bwd(dz1 @ W1, f("norm2.weight"), c_n2)
        put("norm2.weight", gw); put("norm2.bias", gb)
        dx1 = dx1 + d
        put("adapter.scale", np.array([(dy * up).sum()]))
        put("adapter.up_proj.weight", sc * (r2(dy).T @ r2(dn))); put("adapter.up_proj.bias", sc * r2(dy).sum(0))
        dzd = ((dy * sc) @ Wu) * dm3 * (zd > 0)
        put("adapter.down_proj.weight", r2(dzd).T @ r2(an)); put("adapter.down_proj.bias", r2(dzd).sum(0))
        d, gw, gb = _ln_bwd(dzd @ Wd, f("adapter.adapter_norm.weight"), c_an)
        put("adapter.adapter_norm.weight", gw); put("adapter.adapter_norm.bias", gb)
        dx1 = dx1 + d
        datt = dx1 * bc(dp1)
        put("attention.proj.weight", r2(datt).T @ r2(o)); put("attention.proj.bias", r2(datt).sum(0))
        do = (datt @ Wp).reshape(B, G, heads, hd).transpose(0, 2, 1, 3)   # (B,h,G,hd)
        dP = do @ v.transpose(0, 1, 3, 2)
        dv = P.transpose(0, 1, 3, 2) @ do
        ds = P * (dP - (dP * P).sum(-1, keepdims=True)) * hd ** -0.5
        dq = ds @ k
        dk = ds.transpose(0, 1, 3, 2) @ q
        dqkv = np.stack([dq, dk, dv], 0).transpose(1, 3, 0, 2, 4).reshape(B, G, 3 * D)
        put("attention.qkv.weight", r2(dqkv).T @ r2(a)); put("attention.qkv.bias", r2(dqkv).sum(0))
        d, gw, gb = _ln_bwd(dqkv @ Wq, f("norm1.weight"), c_n1)
        put("norm1.weight", gw); put("norm1.bias", gb)
        return dx1 + d

    return y, backward


def apf_vit_backward(sd: Dict[str, np.ndarray], tokens: np.ndarray, depth: int, heads: int, grad_pooled: np.ndarray, masks=None,
                     param_grads: bool = False):
    """Blocks -> encoder_norm -> max over tokens (apf.py:361-366): returns (pooled (B,D), d tokens (B,G,D), gradients) for
    dL/dpooled = grad_pooled.  gradients: encoder_norm.* (trainable in the reference) and, with param_grads, every block
    parameter too (frozen in the reference, apf.py:335-346 - for users who unfreeze them).  masks: per layer, see
    apf_vit_layer_fwd_bwd."""
    x = tokens.astype(np.float64)
    backs = []
    grads: Dict[str, np.ndarray] = {}
    for i in range(depth):
        x, bw = apf_vit_layer_fwd_bwd(sd, f"blocks.{i}.", x, heads, masks[i] if masks is not None else None,
                                      grads if param_grads else None)
        backs.append(bw)
    w, b = np.asarray(sd["encoder_norm.weight"], np.float64), np.asarray(sd["encoder_norm.bias"], np.float64)
    yn, c = _ln_fwd(x, w, b)
    arg = yn.argmax(1)                                                    # (B, D)
    pooled = np.take_along_axis(yn, arg[:, None, :], 1)[:, 0]
    dyn = np.zeros_like(yn)
    np.put_along_axis(dyn, arg[:, None, :], grad_pooled.astype(np.float64)[:, None, :], 1)
    dx, dw, db = _ln_bwd(dyn, w, c)
    for bw in reversed(backs):
        dx = bw(dx)
    grads["encoder_norm.weight"], grads["encoder_norm.bias"] = dw, db
    return pooled, dx, grads


def head_train(sd: Dict[str, np.ndarray], x: np.ndarray, grad_logits: np.ndarray, masks=None, prefix: str = "head.mlp_head."):
    """ClassificationHead in TRAIN mode (apf.py:230-252: Linear, BatchNorm1d with batch statistics, ReLU, Dropout, twice, then
    Linear) and its backward: returns (logits, {"input": dx, parameter gradients by state_dict name}, {running_mean / running_var
    after one momentum update}).  masks = (keep mask (B,512), keep mask (B,256)) scaled by 1 / keep, None = dropout off."""
    f = lambda k: np.asarray(sd[prefix + k], np.float64)
    x = x.astype(np.float64)
    m1, m2 = (1.0 if m is None else np.asarray(m, np.float64) for m in (masks if masks is not None else (None, None)))
    z1 = x @ f("0.weight").T + f("0.bias")
    y1, c1 = _bn_fwd(z1, f("1.weight"), f("1.bias"))
    h1 = np.maximum(y1, 0) * m1
    z2 = h1 @ f("4.weight").T + f("4.bias")
    y2, c2 = _bn_fwd(z2, f("5.weight"), f("5.bias"))
    h2 = np.maximum(y2, 0) * m2
    out = h2 @ f("8.weight").T + f("8.bias")
    g = grad_logits.astype(np.float64)
    grads = {prefix + "8.weight": g.T @ h2, prefix + "8.bias": g.sum(0)}
    dz2, grads[prefix + "5.weight"], grads[prefix + "5.bias"] = _bn_bwd((g @ f("8.weight")) * m2 * (y2 > 0), f("5.weight"), c2)
    grads[prefix + "4.weight"], grads[prefix + "4.bias"] = dz2.T @ h1, dz2.sum(0)
    dz1, grads[prefix + "1.weight"], grads[prefix + "1.bias"] = _bn_bwd((dz2 @ f("4.weight")) * m1 * (y1 > 0), f("1.weight"), c1)
    grads[prefix + "0.weight"], grads[prefix + "0.bias"] = dz1.T @ x, dz1.sum(0)
    grads["input"] = dz1 @ f("0.weight")
    n = x.shape[0]
    running = {}
    for name, (xhat, var, mu) in (("1", c1), ("5", c2)):
        running[prefix + name + ".running_mean"] = (1 - BN_MOMENTUM) * f(name + ".running_mean") + BN_MOMENTUM * mu
        running[prefix + name + ".running_var"] = (1 - BN_MOMENTUM) * f(name + ".running_var") + BN_MOMENTUM * var * n / max(n - 1, 1)
    return out, grads, running


# Backward through Pix4Point's block loop (src/models/pix4point.py:254-271): timm pre-norm Blocks with `feats + pos_embed`
# re-added in front of every block, final norm, 'max,cls' global features.  Pix4Point trains everything (or everything but
# `vit.*` with frozen=True, pix4point.py:229-233), so feats, pos and every parameter receive a gradient.  timm is absent from the
# reference tree (pinned timm==1.0.16): the published Block is restated (oracle.timm_block) and the fixture comes from
# torch.nn.TransformerEncoderLayer(norm_first=True) under autograd (tests/golden/p4p_vit_train.npz).

def pointvit_backward(sd: Dict[str, np.ndarray], feats: np.ndarray, pos: np.ndarray, depth: int, heads: int, grad_glob: np.ndarray,
                      eps: float = 1e-6):
    """-> (global features (B,2D) = [max over tokens 1.. || cls], d feats, d pos, {parameter gradients by state_dict name}) for
    dL/dglobal = grad_glob."""
    x = feats.astype(np.float64)
    pe = pos.astype(np.float64)
    B, S, D = x.shape
    hd = D // heads
    r2 = lambda t: t.reshape(-1, t.shape[-1])
    tape = []
    for i in range(depth):
        p = f"vit.blocks.{i}."
        f = lambda k, p=p: np.asarray(sd[p + k], np.float64)
        xin = x + pe
        a, c1 = _ln_fwd(xin, f("norm1.weight"), f("norm1.bias"), eps)
        qkv = (a @ f("attn.qkv.weight").T + f("attn.qkv.bias")).reshape(B, S, 3, heads, hd).transpose(2, 0, 3, 1, 4)
        q, k, v = qkv[0], qkv[1], qkv[2]
        s = (q @ k.transpose(0, 1, 3, 2)) * hd ** -0.5
        e = np.exp(s - s.max(-1, keepdims=True))
        P = e / e.sum(-1, keepdims=True)
        o = (P @ v).transpose(0, 2, 1, 3).reshape(B, S, D)
        x1 = xin + (o @ f("attn.proj.weight").T + f("attn.proj.bias"))
        n2, c2 = _ln_fwd(x1, f("norm2.weight"), f("norm2.bias"), eps)
        z1 = n2 @ f("mlp.fc1.weight").T + f("mlp.fc1.bias")
        cdf = 0.5 * (1.0 + _erf(z1 / np.sqrt(2.0)))
        h = z1 * cdf
        x = x1 + (h @ f("mlp.fc2.weight").T + f("mlp.fc2.bias"))
        tape.append((p, f, a, c1, q, k, v, P, o, n2, c2, z1, cdf, h))
    w, b = np.asarray(sd["vit.norm.weight"], np.float64), np.asarray(sd["vit.norm.bias"], np.float64)
    yn, cf = _ln_fwd(x, w, b, eps)
    body = yn[:, 1:]
    arg = body.argmax(1)
    glob = np.concatenate([np.take_along_axis(body, arg[:, None, :], 1)[:, 0], yn[:, 0]], 1)
    g = grad_glob.astype(np.float64)
    dyn = np.zeros_like(yn)
    np.put_along_axis(dyn[:, 1:], arg[:, None, :], g[:, None, :D], 1)
    dyn[:, 0] += g[:, D:]
    grads: Dict[str, np.ndarray] = {}
    dx, grads["vit.norm.weight"], grads["vit.norm.bias"] = _ln_bwd(dyn, w, cf)
    dpos = np.zeros_like(pe)
    for (p, f, a, c1, q, k, v, P, o, n2, c2, z1, cdf, h) in reversed(tape):
        grads[p + "mlp.fc2.weight"], grads[p + "mlp.fc2.bias"] = r2(dx).T @ r2(h), r2(dx).sum(0)
        dz1 = (dx @ f("mlp.fc2.weight")) * (cdf + z1 * np.exp(-0.5 * z1 * z1) / np.sqrt(2.0 * np.pi))
        grads[p + "mlp.fc1.weight"], grads[p + "mlp.fc1.bias"] = r2(dz1).T @ r2(n2), r2(dz1).sum(0)
        d, grads[p + "norm2.weight"], grads[p + "norm2.bias"] = _ln_bwd(dz1 @ f("mlp.fc1.weight"), f("norm2.weight"), c2)
        dx1 = dx + d
        grads[p + "attn.proj.weight"], grads[p + "attn.proj.bias"] = r2(dx1).T @ r2(o), r2(dx1).sum(0)
        do = (dx1 @ f("attn.proj.weight")).reshape(B, S, heads, hd).transpose(0, 2, 1, 3)
        dP = do @ v.transpose(0, 1, 3, 2)
        dv = P.transpose(0, 1, 3, 2) @ do
        ds = P * (dP - (dP * P).sum(-1, keepdims=True)) * hd ** -0.5
        dqkv = np.stack([ds @ k, ds.transpose(0, 1, 3, 2) @ q, dv], 0).transpose(1, 3, 0, 2, 4).reshape(B, S, 3 * D)
        grads[p + "attn.qkv.weight"], grads[p + "attn.qkv.bias"] = r2(dqkv).T @ r2(a), r2(dqkv).sum(0)
        d, grads[p + "norm1.weight"], grads[p + "norm1.bias"] = _ln_bwd(dqkv @ f("attn.qkv.weight"), f("norm1.weight"), c1)
        dx = dx1 + d
        dpos = dpos + dx
    return glob, dx, dpos, grads

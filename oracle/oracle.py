"""CPU oracle for the point-patch tokenizer - TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  The product package (p3tok) never does, and raises if its CUDA library is
missing instead of falling back to anything here.

Two layers:
  * index work (FPS, kNN, Morton order, APF grouping): the C restatement in p3tok_oracle.c
    (bit-level spec, see that file's header), called through ctypes;
  * patch embedding (mini-PointNet MLP + max-pool): numpy float64 restatement of the reference
    modules' eval-mode math, layer by layer as written (conv1x1 -> BatchNorm(running stats)
    -> ReLU ...), used as the "true value" the fp32 / bf16 kernels are held to within
    rtol 1e-4 / 1e-2 (BASELINE.json north_star).

Parity status: pinned against outputs of the reference itself (tests/golden/*.npz, generated
in the build container by tests/golden/make_golden.py which imports /root/reference).
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libp3tok_oracle.so")
_lib = None

KNN_APF_SQ = 0
KNN_P4P_CDIST = 1
BN_EPS = 1e-5


def build(force: bool = False) -> str:
    """Compile the C restatement (oracle/Makefile).  Building the checker is not using it."""
    if force or not os.path.isfile(_LIB_PATH) or (
            os.path.getmtime(_LIB_PATH) < os.path.getmtime(os.path.join(_HERE, "p3tok_oracle.c"))):
        subprocess.run(["make", "-C", _HERE, "-s", "-B"], check=True, capture_output=True)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_LIB_PATH)
        i64, fp, ip, vp = ctypes.c_int64, ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_int64), ctypes.c_void_p
        L.orc_fps.argtypes = [fp, i64, i64, i64, ip, i64, ip]
        L.orc_fps_nd.argtypes = [fp, i64, i64, i64, i64, ip, i64, ip]
        L.orc_torch_row_sum.argtypes = [fp, i64, i64, fp]
        L.orc_torch_row_sum.restype = ctypes.c_int
        L.orc_knn.argtypes = [fp, i64, fp, i64, i64, i64, i64, ctypes.c_int, ip, vp]
        L.orc_pair_dist.argtypes = [fp, i64, fp, i64, i64, i64, ctypes.c_int, fp]
        L.orc_morton.argtypes = [fp, i64, i64, ip, vp]
        L.orc_group_apf.argtypes = [fp, i64, i64, i64, ip, ip, vp, i64, i64, fp, fp]
        for f in (L.orc_fps, L.orc_fps_nd, L.orc_knn, L.orc_pair_dist, L.orc_morton, L.orc_group_apf):
            f.restype = ctypes.c_int
        _lib = L
    return _lib


def _f(a: np.ndarray):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))


def _i(a: np.ndarray):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_int64))


def _check(rc: int, what: str):
    if rc != 0:
        raise ValueError(f"oracle {what}: invalid arguments (rc={rc})")


# ----------------------------------------------------------------------------- index work

def fps(x: np.ndarray, start: np.ndarray, G: int) -> np.ndarray:
    """sampler.py:4-30 / pix4point.py:8-53 (without the latter's clamp).  x (B,N,C>=3) float32;
    only channels 0..2 are read.  Returns (B,G) int64."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    B, N, C = x.shape
    start = np.ascontiguousarray(start, dtype=np.int64)
    out = np.empty((B, G), dtype=np.int64)
    _check(lib().orc_fps(_f(x), B, N, C, _i(start), G, _i(out)), "fps")
    return out


def fps_nd(points: np.ndarray, start: np.ndarray, G: int) -> np.ndarray:
    """pix4point.py:8-53 on D-dimensional points (B,N,D), 1 <= D <= 32: the distance sums over ALL D coordinates in
    torch's CPU summation order (p3tok_oracle.c: torch_cpu_row_sum).  No clamp of G.  Returns (B,G) int64."""
    points = np.ascontiguousarray(points, dtype=np.float32)
    B, N, D = points.shape
    start = np.ascontiguousarray(start, dtype=np.int64)
    out = np.empty((B, G), dtype=np.int64)
    _check(lib().orc_fps_nd(_f(points), B, N, D, D, _i(start), G, _i(out)), "fps_nd")
    return out


def torch_row_sum(q: np.ndarray) -> np.ndarray:
    """(rows, D) float32 -> the row sums in the order torch.sum(q, -1) adds them on the CPU (D <= 32)."""
    q = np.ascontiguousarray(q, dtype=np.float32)
    out = np.empty((q.shape[0],), dtype=np.float32)
    _check(lib().orc_torch_row_sum(_f(q), q.shape[0], q.shape[1], _f(out)), "torch_row_sum")
    return out


def knn(x: np.ndarray, centres: np.ndarray, k: int, mode: int, return_dist: bool = False):
    """sampler.py:47-75 (mode KNN_APF_SQ) / pix4point.py:79-89 (mode KNN_P4P_CDIST).
    Canonical result: (B,G,k) int64 ascending by (distance, index)."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    centres = np.ascontiguousarray(centres[..., :3], dtype=np.float32)
    B, N, C = x.shape
    G = centres.shape[1]
    idx = np.empty((B, G, k), dtype=np.int64)
    dist = np.empty((B, G, k), dtype=np.float32) if return_dist else None
    _check(lib().orc_knn(_f(x), C, _f(centres), B, N, G, k, mode, _i(idx),
                         dist.ctypes.data if dist is not None else None), "knn")
    return (idx, dist) if return_dist else idx


def pair_dist(x: np.ndarray, centres: np.ndarray, mode: int) -> np.ndarray:
    x = np.ascontiguousarray(x, dtype=np.float32)
    centres = np.ascontiguousarray(centres[..., :3], dtype=np.float32)
    B, N, C = x.shape
    G = centres.shape[1]
    out = np.empty((B, G, N), dtype=np.float32)
    _check(lib().orc_pair_dist(_f(x), C, _f(centres), B, N, G, mode, _f(out)), "pair_dist")
    return out


def morton(centres: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """apf_utils.py:66-104.  Returns (codes (B,G) int64, stable ascending permutation (B,G))."""
    centres = np.ascontiguousarray(centres, dtype=np.float32)
    B, G, _ = centres.shape
    codes = np.empty((B, G), dtype=np.int64)
    perm = np.empty((B, G), dtype=np.int64)
    _check(lib().orc_morton(_f(centres), B, G, _i(codes), perm.ctypes.data), "morton")
    return codes, perm


def gather_points(x: np.ndarray, idx: np.ndarray) -> np.ndarray:
    """sampler.py:77-94 index_points: x (B,N,C), idx (B,S[,k]) -> (B,S[,k],C)."""
    B = x.shape[0]
    bi = np.arange(B).reshape((B,) + (1,) * (idx.ndim - 1))
    return x[bi, idx]


def group_apf(x: np.ndarray, start: np.ndarray, G: int, k: int, morton_sort: bool = True):
    """apf.py:52-112 Group.forward.  Returns dict(neigh (B,G,k,2C), center (B,G,3), fps_idx,
    knn_idx (canonical order, pre-Morton group order), perm)."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    B, N, C = x.shape
    fidx = fps(x, start, G)
    ctr = gather_points(x[..., :3], fidx)
    kidx = knn(x, ctr, k, KNN_APF_SQ)
    codes, perm = morton(ctr)
    neigh = np.empty((B, G, k, 2 * C), dtype=np.float32)
    center = np.empty((B, G, 3), dtype=np.float32)
    _check(lib().orc_group_apf(_f(x), B, N, C, _i(fidx), _i(kidx),
                               perm.ctypes.data if morton_sort else None, G, k, _f(neigh), _f(center)),
           "group_apf")
    return dict(neigh=neigh, center=center, fps_idx=fidx, knn_idx=kidx, perm=perm, codes=codes)


# ----------------------------------------------------------------------------- embedding (float64)

def _w2(sd: Dict[str, np.ndarray], name: str) -> np.ndarray:
    w = np.asarray(sd[name + ".weight"], dtype=np.float64)
    return w.reshape(w.shape[0], w.shape[1])


def _lin(sd, name, x, bias=True):
    y = x @ _w2(sd, name).T
    if bias and (name + ".bias") in sd:
        y = y + np.asarray(sd[name + ".bias"], dtype=np.float64)
    return y


def _bn(sd, name, x):
    g = np.asarray(sd[name + ".weight"], np.float64)
    b = np.asarray(sd[name + ".bias"], np.float64)
    m = np.asarray(sd[name + ".running_mean"], np.float64)
    v = np.asarray(sd[name + ".running_var"], np.float64)
    return (x - m) / np.sqrt(v + BN_EPS) * g + b


def apf_encoder(sd: Dict[str, np.ndarray], neigh: np.ndarray) -> np.ndarray:
    """apf.py:145-169 Encoder.get_features in eval mode.  neigh (B,G,k,Cin) -> (B,G,E) float64."""
    B, G, k, Cin = neigh.shape
    x = neigh.astype(np.float64).reshape(B * G, k, Cin)
    h = np.maximum(_bn(sd, "first_conv.1", _lin(sd, "first_conv.0", x)), 0)
    h = np.maximum(_bn(sd, "first_conv.4", _lin(sd, "first_conv.3", h)), 0)
    f = _lin(sd, "first_conv.6", h)                       # (BG,k,E)
    g = f.max(axis=1, keepdims=True)                      # apf.py:160
    cat = np.concatenate([np.broadcast_to(g, f.shape), f], -1)   # apf.py:162-163 (global first)
    h = np.maximum(_bn(sd, "second_conv.1", _lin(sd, "second_conv.0", cat)), 0)
    o = _lin(sd, "second_conv.3", h)
    return o.max(axis=1).reshape(B, G, -1)                # apf.py:167-169


def pointnet_apf(sd_encoder: Dict[str, np.ndarray], x: np.ndarray, start: np.ndarray, G: int, k: int):
    """apf.py:202-217 PointNet.forward: Group then Encoder; tokens in Morton group order."""
    grp = group_apf(x, start, G, k, morton_sort=True)
    tok = apf_encoder(sd_encoder, grp["neigh"])
    return tok, grp


def p3embed_stage(sd: Dict[str, np.ndarray], s: int, pts: np.ndarray, feats: np.ndarray,
                  start: np.ndarray, k: int):
    """One iteration of P3Embed.forward's loop (pix4point.py:171-189).  pts (B,N,3) float32,
    feats (B,N,D) (channel-last here), returns (centres (B,N/4,3) f32, tokens (B,N/4,W) f64,
    fps_idx, knn_idx)."""
    B, N, _ = pts.shape
    G = min(N // 4, N)
    fidx = fps(pts, start, G)
    ctr = gather_points(pts, fidx)
    kidx = knn(pts, ctr, k, KNN_P4P_CDIST)
    dp = gather_points(pts, kidx).astype(np.float64)          # absolute coords (pix4point.py:99)
    fj = gather_points(feats, kidx).astype(np.float64)        # (B,G,k,D)
    x = np.concatenate([dp, fj], -1)                          # pix4point.py:182 ([dp, fj])
    p = f"convs.{s}"
    h = _lin(sd, f"{p}.0.0", x, bias=False)                   # no bias / BN / act (135-141)
    h = np.maximum(_bn(sd, f"{p}.0.2", _lin(sd, f"{p}.0.1", h)), 0)
    g = h.max(axis=2, keepdims=True)
    cat = np.concatenate([np.broadcast_to(g, h.shape), h], -1)  # 184-186 (pooled first)
    h2 = np.maximum(_bn(sd, f"{p}.1.1", _lin(sd, f"{p}.1.0", cat, bias=False)), 0)
    h3 = np.maximum(_bn(sd, f"{p}.1.4", _lin(sd, f"{p}.1.3", h2, bias=False)), 0)
    return ctr, h3.max(axis=2), fidx, kidx


def p3embed(sd: Dict[str, np.ndarray], pts: np.ndarray, feats_cl: np.ndarray,
            starts: Sequence[np.ndarray], k: int, stages: int):
    """pix4point.py:166-191 P3Embed.forward.  feats_cl is channel-LAST (B,N,D); the reference
    passes channel-first (B,D,N) and returns channel-first features - callers transpose.
    Stage s>0 consumes the float64 tokens of stage s-1 rounded to float32 (the reference's
    activations are fp32 tensors)."""
    out_p, out_f, aux = [pts], [feats_cl], []
    for s in range(stages):
        ctr, tok, fidx, kidx = p3embed_stage(sd, s, out_p[-1], out_f[-1], starts[s], k)
        out_p.append(ctr)
        out_f.append(tok.astype(np.float32))
        aux.append(dict(fps_idx=fidx, knn_idx=kidx, tokens64=tok))
    return out_p, out_f, aux


def token_head(sd: Dict[str, np.ndarray], tokens: np.ndarray, centres: np.ndarray):
    """pix4point.py:245-252: proj Linear on the tokens, pos_embed MLP (exact erf GELU) on the centres, cls rows first.
    tokens (B,G,W), centres (B,G,3) -> feats, pos (B,1+G,E) float64."""
    from math import sqrt
    try:
        from scipy.special import erf
    except Exception:                                   # pragma: no cover
        erf = np.vectorize(__import__("math").erf)
    f = lambda k: np.asarray(sd[k], np.float64)
    B = tokens.shape[0]
    x = tokens.astype(np.float64) @ f("proj.weight").T + f("proj.bias")
    h = centres.astype(np.float64) @ f("pos_embed.0.weight").T + f("pos_embed.0.bias")
    h = 0.5 * h * (1.0 + erf(h / sqrt(2.0)))
    pe = h @ f("pos_embed.2.weight").T + f("pos_embed.2.bias")
    E = x.shape[-1]
    feats = np.concatenate([np.broadcast_to(f("cls_token").reshape(1, 1, E), (B, 1, E)), x], 1)
    pos = np.concatenate([np.broadcast_to(f("cls_pos").reshape(1, 1, E), (B, 1, E)), pe], 1)
    return feats, pos


def _erf(x):
    try:
        from scipy.special import erf
    except Exception:                                   # pragma: no cover
        erf = np.vectorize(__import__("math").erf)
    return erf(x)


def _layer_norm(sd, name, x, eps=1e-5):
    """nn.LayerNorm over the last axis: biased variance, eps inside the root."""
    mu = x.mean(-1, keepdims=True)
    var = ((x - mu) ** 2).mean(-1, keepdims=True)
    return (x - mu) / np.sqrt(var + eps) * np.asarray(sd[name + ".weight"], np.float64) + np.asarray(sd[name + ".bias"], np.float64)


def apf_vit_layer(sd: Dict[str, np.ndarray], prefix: str, x: np.ndarray, heads: int) -> np.ndarray:
    """One APFViTLayer in eval mode, float64 (src/models/apf_utils.py:268-293): x (B,G,D) -> (B,G,D).
    attention = AttentionLayer.forward (apf_utils.py:133-160), adapter = AdapterLayer.forward (197-233; returns
    scale*up + x), mlp = fc1 -> exact GELU -> fc2; out = mlp(norm2(x)) + adapter(x) + x (line 292)."""
    f = lambda k: np.asarray(sd[prefix + k], np.float64)
    B, G, D = x.shape
    hd = D // heads
    a = _layer_norm(sd, prefix + "norm1", x)
    qkv = (a @ f("attention.qkv.weight").T + f("attention.qkv.bias")).reshape(B, G, 3, heads, hd).transpose(2, 0, 3, 1, 4)
    q, k, v = qkv[0], qkv[1], qkv[2]
    att = (q @ k.transpose(0, 1, 3, 2)) * hd ** -0.5
    att = np.exp(att - att.max(-1, keepdims=True))
    att = att / att.sum(-1, keepdims=True)
    o = (att @ v).transpose(0, 2, 1, 3).reshape(B, G, D)
    x = x + (o @ f("attention.proj.weight").T + f("attention.proj.bias"))
    an = _layer_norm(sd, prefix + "adapter.adapter_norm", x)
    down = np.maximum(an @ f("adapter.down_proj.weight").T + f("adapter.down_proj.bias"), 0.0)
    adapt = (down @ f("adapter.up_proj.weight").T + f("adapter.up_proj.bias")) * float(np.asarray(sd[prefix + "adapter.scale"]).reshape(-1)[0]) + x
    h = _layer_norm(sd, prefix + "norm2", x) @ f("mlp.fc1.weight").T + f("mlp.fc1.bias")
    h = 0.5 * h * (1.0 + _erf(h / np.sqrt(2.0)))
    return (h @ f("mlp.fc2.weight").T + f("mlp.fc2.bias")) + adapt + x


def apf_vit(sd: Dict[str, np.ndarray], tokens: np.ndarray, depth: int, heads: int):
    """Block loop + encoder_norm + max over tokens + classification head of AdaptPointFormer.forward
    (src/models/apf.py:361-371), eval mode, float64.  Returns (x (B,G,D), pooled (B,D), logits (B,classes))."""
    x = tokens.astype(np.float64)
    for i in range(depth):
        x = apf_vit_layer(sd, f"blocks.{i}.", x, heads)
    pooled = _layer_norm(sd, "encoder_norm", x).max(1)
    f = lambda k: np.asarray(sd[k], np.float64)
    h = pooled
    for lin, bn in (("head.mlp_head.0", "head.mlp_head.1"), ("head.mlp_head.4", "head.mlp_head.5")):
        h = h @ f(lin + ".weight").T + f(lin + ".bias")
        h = (h - f(bn + ".running_mean")) / np.sqrt(f(bn + ".running_var") + BN_EPS) * f(bn + ".weight") + f(bn + ".bias")
        h = np.maximum(h, 0.0)
    logits = h @ f("head.mlp_head.8.weight").T + f("head.mlp_head.8.bias")
    return x, pooled, logits


def timm_block(sd: Dict[str, np.ndarray], prefix: str, x: np.ndarray, heads: int, eps: float = 1e-6) -> np.ndarray:
    """One pre-norm ViT block as Pix4Point runs it (src/models/pix4point.py:254-255 calls timm's `Block`; timm is absent from
    the reference tree - pinned timm==1.0.16 in requirements.txt - so its published forward is restated:
    x = x + proj(softmax(q k^T / sqrt(hd)) v) on norm1(x);  x = x + fc2(gelu(fc1(norm2(x)))), LayerNorm eps 1e-6, qkv with
    bias, no LayerScale (`init_values=None` for vit_small_patch16_384), DropPath / dropout = identity in eval), float64."""
    f = lambda k: np.asarray(sd[prefix + k], np.float64)
    B, S, D = x.shape
    hd = D // heads
    a = _layer_norm(sd, prefix + "norm1", x, eps)
    qkv = (a @ f("attn.qkv.weight").T + f("attn.qkv.bias")).reshape(B, S, 3, heads, hd).transpose(2, 0, 3, 1, 4)
    q, k, v = qkv[0], qkv[1], qkv[2]
    att = (q @ k.transpose(0, 1, 3, 2)) * hd ** -0.5
    att = np.exp(att - att.max(-1, keepdims=True))
    att = att / att.sum(-1, keepdims=True)
    o = (att @ v).transpose(0, 2, 1, 3).reshape(B, S, D)
    x = x + (o @ f("attn.proj.weight").T + f("attn.proj.bias"))
    h = _layer_norm(sd, prefix + "norm2", x, eps) @ f("mlp.fc1.weight").T + f("mlp.fc1.bias")
    h = 0.5 * h * (1.0 + _erf(h / np.sqrt(2.0)))
    return x + (h @ f("mlp.fc2.weight").T + f("mlp.fc2.bias"))


def pointvit_blocks(sd: Dict[str, np.ndarray], feats: np.ndarray, pos: np.ndarray, depth: int, heads: int, eps: float = 1e-6):
    """The tail of PointViT.forward (pix4point.py:254-256) + forward_cls_feat (260-271, global_features 'max,cls'):
    feats, pos (B,1+G,D) -> (normed (B,1+G,D), global (B,2D) = [max over the tokens without cls || cls])."""
    x = feats.astype(np.float64)
    p = pos.astype(np.float64)
    for i in range(depth):
        x = timm_block(sd, f"vit.blocks.{i}.", x + p, heads, eps)
    x = _layer_norm(sd, "vit.norm", x, eps)
    return x, np.concatenate([x[:, 1:].max(1), x[:, 0]], 1)


# ----------------------------------------------------------------------------- comparison helpers

def knn_tie_equivalent(idx_a: np.ndarray, idx_b: np.ndarray, dist_full: np.ndarray,
                       ulp: int = 0) -> Tuple[bool, str]:
    """True when two (B,G,k) neighbour index sets are equal up to ties at the k-th distance:
    every index that is in one set and not in the other must have a distance within `ulp`
    units-in-the-last-place of that group's k-th smallest distance (SURVEY.md hard part 2)."""
    B, G, k = idx_a.shape
    a = np.sort(idx_a.reshape(B * G, k), -1)
    b = np.sort(idx_b.reshape(B * G, k), -1)
    rows = np.nonzero((a != b).any(-1))[0]
    D = dist_full.reshape(B * G, -1)
    for r in rows:
        sa, sb = set(a[r].tolist()), set(b[r].tolist())
        if len(sa) != k or len(sb) != k:
            return False, f"group {r}: duplicate indices"
        kth = np.sort(D[r])[k - 1]
        for i in (sa ^ sb):
            if _ulp_diff(D[r, i], kth) > ulp:
                return False, f"group {r}: index {i} dist {D[r, i]!r} vs kth {kth!r}"
    return True, f"{len(rows)} tied groups of {B * G}"


def _ulp_diff(x, y) -> int:
    xi = int(np.float32(x).view(np.int32))
    yi = int(np.float32(y).view(np.int32))
    xi = xi if xi >= 0 else -(xi & 0x7FFFFFFF)
    yi = yi if yi >= 0 else -(yi & 0x7FFFFFFF)
    return abs(xi - yi)


def sorted_by_distance(idx: np.ndarray, dist_full: np.ndarray, ulp: int = 0) -> bool:
    """Order check for sorted=True outputs: distances along k never decrease by more than ulp."""
    B, G, k = idx.shape
    D = np.take_along_axis(dist_full, idx.astype(np.int64), axis=-1)
    di = D.view(np.int32).astype(np.int64)      # distances are >= 0 here -> int order == float order
    return bool((di[..., 1:] - di[..., :-1] >= -ulp).all())

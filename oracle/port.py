"""torch-CPU port of the reference tokenizer - TEST / BASELINE INFRASTRUCTURE, NOT PRODUCT CODE.

This is the `cpu_baseline` (kind "port") that bench.py times on the GPU box's host cores and
the arm `bench.py --impl reference` runs: the reference is pure Python and `/root/reference`
does not travel to the GPU box, so its tokenizer is restated here with the SAME torch CPU
calls in the SAME order (so it costs what the reference costs and, on one host, returns the
same bits - asserted against the imported reference by tests/test_oracle_vs_reference.py in
the build container).  The only intended difference: the FPS start index is an argument
instead of an internal torch.randint draw (sampler.py:20, pix4point.py:30).

Functional style with explicit state_dicts; no nn.Module classes of the reference are reused.
"""
from __future__ import annotations

from typing import Dict, List, Sequence, Tuple

import torch
import torch.nn.functional as F

BN_EPS = 1e-5


# --- point ops ---------------------------------------------------------------------------------

def fps_indices(pts: torch.Tensor, n: int, start: torch.Tensor, masked_update: bool = False) -> torch.Tensor:
    """sampler.py:4-30 (masked_update=False: torch.min) / pix4point.py:8-53 (True: mask assign)."""
    B, N, _ = pts.shape
    rows = torch.arange(B)
    picked = torch.zeros(B, n, dtype=torch.long)
    closest = torch.full((B, N), 1e10)
    cur = start.clone().long()
    for j in range(n):
        picked[:, j] = cur
        c = pts[rows, cur, :].view(B, 1, -1)
        d = torch.sum((pts - c) ** 2, -1)
        if masked_update:
            m = d < closest
            closest[m] = d[m]
        else:
            closest = torch.min(closest, d)
        cur = torch.max(closest, -1)[1]
    return picked


def sq_dist_expanded(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """sampler.py:47-62: -2ab^T + |a|^2 + |b|^2, accumulated in that order."""
    B, S, _ = a.shape
    N = b.shape[1]
    d = -2 * torch.matmul(a, b.permute(0, 2, 1))
    d += torch.sum(a ** 2, -1).view(B, S, 1)
    d += torch.sum(b ** 2, -1).view(B, 1, N)
    return d


def knn_apf(k: int, pts: torch.Tensor, queries: torch.Tensor) -> torch.Tensor:
    """sampler.py:64-75: unordered k smallest of the expanded squared distance."""
    return torch.topk(sq_dist_expanded(queries, pts), k, dim=-1, largest=False, sorted=False)[1]


def knn_p4p(k: int, pts: torch.Tensor, queries: torch.Tensor) -> torch.Tensor:
    """pix4point.py:79-89: cdist + sorted topk, int32 indices."""
    return torch.cdist(queries, pts).topk(k=k, dim=-1, largest=False, sorted=True).indices.int()


def take(points: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
    """sampler.py:77-94 index_points."""
    B = points.shape[0]
    shape = [B] + [1] * (idx.dim() - 1)
    return points[torch.arange(B).view(shape).expand_as(idx), idx.long(), :]


def morton_order(c: torch.Tensor, resolution: int = 1024) -> torch.Tensor:
    """apf_utils.py:34-104: argsort of the 10-bit/axis Z-order code of the centres."""
    lo = c.min(dim=1, keepdim=True)[0]
    hi = c.max(dim=1, keepdim=True)[0]
    q = (((c - lo) / (hi - lo + 1e-8)) * (resolution - 1)).long()

    def spread(n):
        n = n & 0x000003ff
        n = (n ^ (n << 16)) & 0xff0000ff
        n = (n ^ (n << 8)) & 0x0300f00f
        n = (n ^ (n << 4)) & 0x030c30c3
        n = (n ^ (n << 2)) & 0x09249249
        return n

    code = (spread(q[..., 2]) << 2) + (spread(q[..., 1]) << 1) + spread(q[..., 0])
    return torch.argsort(code, dim=1)


# --- APF tokenizer -----------------------------------------------------------------------------

def apf_group(x: torch.Tensor, G: int, k: int, start: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """apf.py:52-112 Group.forward incl. the dead cdist(center,center)+eye-mask work (apf.py:38-39)
    the reference pays for."""
    B, N, _ = x.shape
    xyz = x[:, :, :3].contiguous()
    fidx = fps_indices(xyz, G, start)
    center = take(xyz, fidx)
    cfeat = take(x, fidx)
    idx = knn_apf(k, xyz, center)
    flat = (idx + torch.arange(B).view(-1, 1, 1) * N).view(-1)
    neigh = x.view(B * N, -1)[flat, :].view(B, G, k, -1).contiguous()
    neigh = neigh - cfeat.unsqueeze(-2)
    neigh = torch.cat([neigh, cfeat.unsqueeze(2).repeat(1, 1, k, 1)], dim=-1)
    dead = torch.cdist(center, center)
    dead[:, torch.eye(G).bool()] = float("inf")
    order = (morton_order(center) + (torch.arange(B) * G).unsqueeze(1)).view(-1)
    neigh = neigh.view(B * G, k, -1)[order].view(B, G, k, -1).contiguous()
    center = center.view(B * G, -1)[order].view(B, G, -1).contiguous()
    return neigh, center


def _conv_bn_relu_1d(sd, conv: str, bn: str, h: torch.Tensor) -> torch.Tensor:
    h = F.conv1d(h, sd[conv + ".weight"], sd.get(conv + ".bias"))
    h = F.batch_norm(h, sd[bn + ".running_mean"], sd[bn + ".running_var"], sd[bn + ".weight"],
                     sd[bn + ".bias"], False, 0.1, BN_EPS)
    return F.relu(h, inplace=True)


def apf_encode(sd: Dict[str, torch.Tensor], groups: torch.Tensor) -> torch.Tensor:
    """apf.py:145-169 Encoder.get_features (eval)."""
    B, G, k, _ = groups.shape
    h = groups.reshape(B * G, k, -1).transpose(2, 1)
    h = _conv_bn_relu_1d(sd, "first_conv.0", "first_conv.1", h)
    h = _conv_bn_relu_1d(sd, "first_conv.3", "first_conv.4", h)
    h = F.conv1d(h, sd["first_conv.6.weight"], sd["first_conv.6.bias"])
    g = torch.max(h, dim=2, keepdim=True)[0]
    h = torch.cat([g.expand(-1, -1, k), h], dim=1)
    h = _conv_bn_relu_1d(sd, "second_conv.0", "second_conv.1", h)
    h = F.conv1d(h, sd["second_conv.3.weight"], sd["second_conv.3.bias"])
    return torch.max(h, dim=2, keepdim=False)[0].reshape(B, G, -1)


def apf_pointnet(sd: Dict[str, torch.Tensor], x: torch.Tensor, G: int, k: int, start: torch.Tensor) -> torch.Tensor:
    """apf.py:202-217 PointNet.forward."""
    neigh, _ = apf_group(x, G, k, start)
    return apf_encode(sd, neigh)


# --- APF token consumer ----------------------------------------------------

def apf_vit_layer(sd: Dict[str, torch.Tensor], p: str, x: torch.Tensor, heads: int) -> torch.Tensor:
    """apf_utils.py:268-293 APFViTLayer.forward (eval: DropPath / dropout are the identity), the reference's torch calls in
    its order: norm1 -> AttentionLayer (133-160) -> residual; AdapterLayer (197-233); norm2 -> Mlp -> sum."""
    B, N, C = x.shape
    a = F.layer_norm(x, (C,), sd[p + "norm1.weight"], sd[p + "norm1.bias"], 1e-5)
    qkv = F.linear(a, sd[p + "attention.qkv.weight"], sd[p + "attention.qkv.bias"])
    qkv = qkv.reshape(B, N, 3, heads, C // heads).permute(2, 0, 3, 1, 4)
    q, k, v = qkv[0], qkv[1], qkv[2]
    attn = (q @ k.transpose(-2, -1)) * (C // heads) ** -0.5
    attn = attn.softmax(dim=-1)
    o = (attn @ v).transpose(1, 2).reshape(B, N, C)
    x = x + F.linear(o, sd[p + "attention.proj.weight"], sd[p + "attention.proj.bias"])
    residual = x
    an = F.layer_norm(x, (C,), sd[p + "adapter.adapter_norm.weight"], sd[p + "adapter.adapter_norm.bias"], 1e-5)
    down = F.relu(F.linear(an, sd[p + "adapter.down_proj.weight"], sd[p + "adapter.down_proj.bias"]))
    up = F.linear(down, sd[p + "adapter.up_proj.weight"], sd[p + "adapter.up_proj.bias"]) * sd[p + "adapter.scale"]
    adapt = up + x
    m = F.layer_norm(x, (C,), sd[p + "norm2.weight"], sd[p + "norm2.bias"], 1e-5)
    m = F.linear(F.gelu(F.linear(m, sd[p + "mlp.fc1.weight"], sd[p + "mlp.fc1.bias"])), sd[p + "mlp.fc2.weight"],
                 sd[p + "mlp.fc2.bias"])
    return m + adapt + residual


def apf_vit_features(sd: Dict[str, torch.Tensor], tokens: torch.Tensor, depth: int, heads: int) -> torch.Tensor:
    """apf.py:361-366: the block loop, encoder_norm and the max over tokens -> (B, D)."""
    x = tokens
    for i in range(depth):
        x = apf_vit_layer(sd, f"blocks.{i}.", x, heads)
    C = x.shape[-1]
    x = F.layer_norm(x, (C,), sd["encoder_norm.weight"], sd["encoder_norm.bias"], 1e-5)
    return x.max(-2)[0]


# --- Pix4Point tokenizer -----------------------------------------------------------------------

def _bn2(sd, bn: str, h: torch.Tensor) -> torch.Tensor:
    return F.batch_norm(h, sd[bn + ".running_mean"], sd[bn + ".running_var"], sd[bn + ".weight"],
                        sd[bn + ".bias"], False, 0.1, BN_EPS)


def p3embed(sd: Dict[str, torch.Tensor], p: torch.Tensor, f: torch.Tensor, k: int, stages: int,
            starts: Sequence[torch.Tensor]) -> Tuple[List[torch.Tensor], List[torch.Tensor]]:
    """pix4point.py:166-191 P3Embed.forward (layers=4).  f is channel-first (B,D,N) as in the reference."""
    B, N, _ = p.shape
    ps, fs = [p], [f]
    for s in range(stages):
        pts, feat = ps[-1], fs[-1].transpose(1, 2)
        N = N // 4
        cidx = fps_indices(pts, min(N, pts.shape[1]), starts[s], masked_update=True)
        ctr = torch.gather(pts, 1, cidx.unsqueeze(-1).expand(-1, -1, 3))
        nidx = knn_p4p(k, pts, ctr)
        bi = torch.arange(B).view(B, 1, 1).expand(-1, ctr.shape[1], k)
        dp = pts[bi, nidx].permute(0, 3, 1, 2).contiguous()
        fj = feat[bi, nidx].permute(0, 3, 1, 2).contiguous()
        h = torch.cat([dp, fj], dim=1)
        pre = f"convs.{s}"
        h = F.conv2d(h, sd[f"{pre}.0.0.weight"])
        h = F.relu(_bn2(sd, f"{pre}.0.2", F.conv2d(h, sd[f"{pre}.0.1.weight"], sd[f"{pre}.0.1.bias"])))
        h = torch.cat([torch.max(h, dim=-1, keepdim=True)[0].expand(-1, -1, -1, k), h], dim=1)
        h = F.relu(_bn2(sd, f"{pre}.1.1", F.conv2d(h, sd[f"{pre}.1.0.weight"])))
        h = F.relu(_bn2(sd, f"{pre}.1.4", F.conv2d(h, sd[f"{pre}.1.3.weight"])))
        fs.append(torch.max(h, dim=-1, keepdim=True)[0].squeeze(-1))
        ps.append(ctr)
    return ps, fs

"""Deterministic synthetic clouds and tokenizer weights.

Everything here is a pure function of (seed, shape) built from a counter-based integer
hash evaluated with numpy uint64 arithmetic, so the build container, the GPU box, the
golden-vector generator and bench.py all see bit-identical inputs without shipping
tensors around (torch's global RNG - what the reference draws FPS start indices from,
src/data/sampler.py:20 - differs between CPU and CUDA generators).

Cloud kinds follow SURVEY.md 8d: "uniform" in [-1,1)^3, "clustered" (8 blobs, centres in
U(-0.8,0.8)^3, sigma ~= 0.05) and "duplicates" (uniform points sampled with replacement,
the tie-stress case that mirrors src/data/scanobjectnn.py:171-181).
"""
from __future__ import annotations

import math
from typing import Dict, Tuple

import numpy as np

_M1 = np.uint64(0xBF58476D1CE4E5B9)
_M2 = np.uint64(0x94D049BB133111EB)
_G = np.uint64(0x9E3779B97F4A7C15)


def _mix(z: np.ndarray) -> np.ndarray:
    z = z.astype(np.uint64, copy=True)
    with np.errstate(over="ignore"):
        z ^= z >> np.uint64(30)
        z *= _M1
        z ^= z >> np.uint64(27)
        z *= _M2
        z ^= z >> np.uint64(31)
    return z


def hash_u64(seed: int, n: int, stream: int = 0) -> np.ndarray:
    """n 64-bit hashes for counters 0..n-1 of (seed, stream)."""
    with np.errstate(over="ignore"):
        base = _mix(np.array([np.uint64(seed & 0xFFFFFFFFFFFFFFFF) * _G
                              + np.uint64(stream) * _M2 + np.uint64(0x1234567)], dtype=np.uint64))[0]
        i = np.arange(n, dtype=np.uint64)
        return _mix(_mix(i * _G + base) + base)


def uniform01(seed: int, n: int, stream: int = 0) -> np.ndarray:
    """float32 uniforms m/2^24 in [0,1) - exactly representable, platform independent."""
    return (hash_u64(seed, n, stream) >> np.uint64(40)).astype(np.float32) * np.float32(2.0 ** -24)


def randint(seed: int, n: int, high: int, stream: int = 0) -> np.ndarray:
    return (hash_u64(seed, n, stream) % np.uint64(high)).astype(np.int64)


def make_cloud(kind: str, B: int, N: int, seed: int, channels: int = 3) -> np.ndarray:
    """(B,N,channels) float32.  channels=4 appends the 'height' feature y - min(y)
    (src/data/augment.py:247-248) as the 4th channel."""
    if kind == "uniform":
        xyz = (uniform01(seed, B * N * 3, 1) * np.float32(2) - np.float32(1)).reshape(B, N, 3)
    elif kind == "clustered":
        nb = 8
        ctr = (uniform01(seed, B * nb * 3, 2).astype(np.float64) * 1.6 - 0.8).reshape(B, nb, 3)
        which = randint(seed, B * N, nb, 3).reshape(B, N)
        u = uniform01(seed, B * N * 3 * 4, 4).astype(np.float64).reshape(B, N, 3, 4)
        # Irwin-Hall(4): mean 2, var 1/3 -> unit variance after *sqrt(3); exact in fp64
        noise = (u.sum(-1) - 2.0) * math.sqrt(3.0) * 0.05
        xyz = (np.take_along_axis(ctr, which[..., None].repeat(3, -1), axis=1) + noise).astype(np.float32)
    elif kind == "duplicates":
        base = make_cloud("uniform", B, N, seed, 3)
        pick = randint(seed, B * N, N, 5).reshape(B, N)
        xyz = np.take_along_axis(base, pick[..., None].repeat(3, -1), axis=1)
    else:
        raise ValueError(f"unknown cloud kind {kind!r}")
    xyz = np.ascontiguousarray(xyz, dtype=np.float32)
    if channels == 3:
        return xyz
    if channels == 4:
        h = xyz[..., 1:2] - xyz[..., 1:2].min(axis=1, keepdims=True)
        return np.ascontiguousarray(np.concatenate([xyz, h], -1), dtype=np.float32)
    raise ValueError("channels must be 3 or 4")


def make_points_nd(B: int, N: int, D: int, seed: int) -> np.ndarray:
    """(B,N,D) float32 points uniform in [-1,1)^D for the general-D FPS (farthest_point_sampling, pix4point.py:8-53);
    the second quarter of every cloud repeats the first, so exact distance ties (lowest index wins) are exercised."""
    pts = (uniform01(seed, B * N * D, 7 + D) * np.float32(2) - np.float32(1)).reshape(B, N, D)
    q = N // 4
    pts[:, q:2 * q] = pts[:, :q]
    return np.ascontiguousarray(pts, dtype=np.float32)


def start_indices(B: int, N: int, seed: int, stage: int = 0) -> np.ndarray:
    """FPS start index per cloud (stands in for torch.randint(0,N,(B,)), sampler.py:20)."""
    return randint(seed, B, N, 100 + stage)


class _WeightStream:
    def __init__(self, seed: int):
        self.seed = seed
        self.stream = 1000

    def uniform(self, shape, lo: float, hi: float) -> np.ndarray:
        n = int(np.prod(shape))
        self.stream += 1
        u = uniform01(self.seed, n, self.stream).astype(np.float64)
        return (lo + (hi - lo) * u).astype(np.float32).reshape(shape)


def _conv(ws: _WeightStream, sd: Dict[str, np.ndarray], name: str, cout: int, cin: int,
          bias: bool, ndim: int) -> None:
    bound = 1.0 / math.sqrt(cin)
    shape = (cout, cin) + (1,) * (ndim - 2)
    sd[name + ".weight"] = ws.uniform(shape, -bound, bound)
    if bias:
        sd[name + ".bias"] = ws.uniform((cout,), -bound, bound)


def _bn(ws: _WeightStream, sd: Dict[str, np.ndarray], name: str, c: int) -> None:
    sd[name + ".weight"] = ws.uniform((c,), 0.5, 1.5)
    sd[name + ".bias"] = ws.uniform((c,), -0.1, 0.1)
    sd[name + ".running_mean"] = ws.uniform((c,), -0.1, 0.1)
    sd[name + ".running_var"] = ws.uniform((c,), 0.5, 1.5)
    sd[name + ".num_batches_tracked"] = np.array(0, dtype=np.int64)


def apf_encoder_state(embed_dim: int, in_channel: int, seed: int = 0) -> Dict[str, np.ndarray]:
    """state_dict (numpy) with the reference Encoder's keys (src/models/apf.py:129-143):
    first_conv.{0,1,3,4,6}.*, second_conv.{0,1,3}.*; BN running stats non-trivial."""
    ws = _WeightStream(seed)
    sd: Dict[str, np.ndarray] = {}
    E = embed_dim
    _conv(ws, sd, "first_conv.0", 256, in_channel, True, 3)
    _bn(ws, sd, "first_conv.1", 256)
    _conv(ws, sd, "first_conv.3", 512, 256, True, 3)
    _bn(ws, sd, "first_conv.4", 512)
    _conv(ws, sd, "first_conv.6", E, 512, True, 3)
    _conv(ws, sd, "second_conv.0", 2 * E, 2 * E, True, 3)
    _bn(ws, sd, "second_conv.1", 2 * E)
    _conv(ws, sd, "second_conv.3", E, 2 * E, True, 3)
    return sd


def p3embed_dims(in_channels: int = 3, sample_ratio: float = 0.25, scale: int = 4,
                 layers: int = 4, embed_dim: int = 256) -> Tuple[int, list]:
    """(stages, [(Cin, W), ...]) exactly as P3Embed.__init__ derives them
    (src/models/pix4point.py:123-160)."""
    stages = int(math.log(1 / sample_ratio, scale))
    w = int(embed_dim // 2 ** (stages - 1))
    dims = []
    cin = in_channels
    for _ in range(stages):
        dims.append((cin + 3, w))
        cin = w
        w *= 2
    return stages, dims


def p3embed_state(in_channels: int = 3, sample_ratio: float = 0.25, scale: int = 4,
                  layers: int = 4, embed_dim: int = 256, seed: int = 0) -> Dict[str, np.ndarray]:
    """state_dict (numpy) with P3Embed's keys for layers=4: convs.{s}.0.{0,1,2}.*, convs.{s}.1.{0,1,3,4}.*"""
    if layers != 4:
        raise ValueError("only the reference's layers=4 layout is supported")
    ws = _WeightStream(seed)
    sd: Dict[str, np.ndarray] = {}
    _, dims = p3embed_dims(in_channels, sample_ratio, scale, layers, embed_dim)
    for s, (cin, w) in enumerate(dims):
        p = f"convs.{s}"
        _conv(ws, sd, f"{p}.0.0", w, cin, False, 4)
        _conv(ws, sd, f"{p}.0.1", w, w, True, 4)
        _bn(ws, sd, f"{p}.0.2", w)
        _conv(ws, sd, f"{p}.1.0", 2 * w, 2 * w, False, 4)
        _bn(ws, sd, f"{p}.1.1", 2 * w)
        _conv(ws, sd, f"{p}.1.3", w, 2 * w, False, 4)
        _bn(ws, sd, f"{p}.1.4", w)
    return sd


def token_head_state(width: int, embed_dim: int, seed: int = 0) -> Dict[str, np.ndarray]:
    """proj / pos_embed / cls parameters of the reference's PointViT (src/models/pix4point.py:213-218, 229-230)."""
    ws = _WeightStream(seed + 7919)
    sd: Dict[str, np.ndarray] = {}
    for name, cout, cin in (("proj", embed_dim, width), ("pos_embed.0", 128, 3), ("pos_embed.2", embed_dim, 128)):
        bound = 1.0 / math.sqrt(cin)
        sd[name + ".weight"] = ws.uniform((cout, cin), -bound, bound)
        sd[name + ".bias"] = ws.uniform((cout,), -bound, bound)
    sd["cls_token"] = ws.uniform((1, 1, embed_dim), -0.02, 0.02)
    sd["cls_pos"] = ws.uniform((1, 1, embed_dim), -0.02, 0.02)
    return sd


def apf_vit_state(dim: int, depth: int, num_classes: int = 15, seed: int = 0, bottleneck: int = 64) -> Dict[str, np.ndarray]:
    """state_dict (numpy) of the token consumer of the reference's AdaptPointFormer (src/models/apf.py:296-317): `blocks.{i}.*`
    (APFViTLayer, src/models/apf_utils.py:236-266), `encoder_norm.*`, `head.mlp_head.*`.  Unlike the reference's initialisation
    the adapter's up-projection and scale are non-trivial, so every term of the layer is exercised."""
    ws = _WeightStream(seed + 104729)
    sd: Dict[str, np.ndarray] = {}

    def lin(name, cout, cin, scale=1.0):
        bound = scale / math.sqrt(cin)
        sd[name + ".weight"] = ws.uniform((cout, cin), -bound, bound)
        sd[name + ".bias"] = ws.uniform((cout,), -bound, bound)

    def ln(name):
        sd[name + ".weight"] = ws.uniform((dim,), 0.5, 1.5)
        sd[name + ".bias"] = ws.uniform((dim,), -0.1, 0.1)

    for i in range(depth):
        b = f"blocks.{i}."
        ln(b + "norm1")
        ln(b + "norm2")
        lin(b + "mlp.fc1", 4 * dim, dim)
        lin(b + "mlp.fc2", dim, 4 * dim)
        lin(b + "attention.qkv", 3 * dim, dim, 2.0)
        lin(b + "attention.proj", dim, dim)
        ln(b + "adapter.adapter_norm")
        sd[b + "adapter.scale"] = ws.uniform((1,), 0.5, 0.9)
        lin(b + "adapter.down_proj", bottleneck, dim)
        lin(b + "adapter.up_proj", dim, bottleneck)
    ln("encoder_norm")
    lin("head.mlp_head.0", 512, dim)
    _bn(ws, sd, "head.mlp_head.1", 512)
    lin("head.mlp_head.4", 256, 512)
    _bn(ws, sd, "head.mlp_head.5", 256)
    lin("head.mlp_head.8", num_classes, 256)
    return sd


def pointvit_state(dim: int, depth: int, seed: int = 0, mlp_ratio: int = 4) -> Dict[str, np.ndarray]:
    """state_dict (numpy) of the timm ViT inside the reference's PointViT (src/models/pix4point.py:221-228): `vit.blocks.{i}.*`
    with timm's Block layout (norm1, attn.qkv, attn.proj, norm2, mlp.fc1, mlp.fc2) and `vit.norm.*`."""
    ws = _WeightStream(seed + 15485863)
    sd: Dict[str, np.ndarray] = {}

    def lin(name, cout, cin, scale=1.0):
        bound = scale / math.sqrt(cin)
        sd[name + ".weight"] = ws.uniform((cout, cin), -bound, bound)
        sd[name + ".bias"] = ws.uniform((cout,), -bound, bound)

    def ln(name):
        sd[name + ".weight"] = ws.uniform((dim,), 0.5, 1.5)
        sd[name + ".bias"] = ws.uniform((dim,), -0.1, 0.1)

    for i in range(depth):
        b = f"vit.blocks.{i}."
        ln(b + "norm1")
        lin(b + "attn.qkv", 3 * dim, dim, 2.0)
        lin(b + "attn.proj", dim, dim)
        ln(b + "norm2")
        lin(b + "mlp.fc1", mlp_ratio * dim, dim)
        lin(b + "mlp.fc2", dim, mlp_ratio * dim)
    ln("vit.norm")
    return sd


def vit_tokens(B: int, G: int, D: int, seed: int) -> np.ndarray:
    """Synthetic (B,G,D) token batch in [-1,1) for the ViT block-stack cases."""
    u = uniform01(seed, B * G * D, 17).reshape(B, G, D)
    return ((u - 0.5) * 2.0).astype(np.float32)


def to_torch_state(sd: Dict[str, np.ndarray]):
    import torch
    return {k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in sd.items()}

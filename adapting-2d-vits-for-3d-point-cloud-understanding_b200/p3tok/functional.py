"""Free functions with the reference's signatures, backed by the sm_100a kernels.

Drop-in for `from data.sampler import furthest_point_sample, fps, knn_point, index_points`
(reference src/data/sampler.py:4-94) and for `farthest_point_sampling`, `group_knn`
(src/models/pix4point.py:8-102).  Every function takes one optional extra keyword the
reference does not have - `start_idx` - because the reference draws the first FPS index from
torch's global RNG inside the call (sampler.py:20).  When it is omitted the same draw is made
here the way each reference function makes it, so `torch.manual_seed(s)` reproduces the
reference's centres: furthest_point_sample / fps draw `torch.randint` on the global CPU
generator and copy the result to the device (sampler.py:20); farthest_point_sampling draws on
the input's device (pix4point.py:30).  (A host draw cannot be captured into a CUDA graph: pass
`start_idx` when capturing, as p3tok.graph does.)
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from . import _lib, ops


def _start(x: torch.Tensor, start_idx: Optional[torch.Tensor], device_draw: bool = False) -> torch.Tensor:
    B, N = x.shape[0], x.shape[1]
    if start_idx is None:
        if device_draw:
            return torch.randint(0, N, (B,), dtype=torch.long, device=x.device)   # pix4point.py:30
        return torch.randint(0, N, (B,), dtype=torch.long).to(x.device)           # sampler.py:20: CPU generator, then .to(device)
    return start_idx.to(device=x.device, dtype=torch.long)


def furthest_point_sample(xyz: torch.Tensor, npoint: int, start_idx: Optional[torch.Tensor] = None) -> torch.Tensor:
    """sampler.py:4-30: (B,N,3) -> (B,npoint) int64; column 0 is the start index; no clamp."""
    return ops.fps(xyz, _start(xyz, start_idx), int(npoint))


def farthest_point_sampling(points: torch.Tensor, n_samples: int,
                            start_idx: Optional[torch.Tensor] = None) -> torch.Tensor:
    """pix4point.py:8-53: clamps n_samples to N (line 23); distances sum over ALL D coordinates (line 44).
    D = 3 - the reference's only call site passes xyz (pix4point.py:175) - runs the register / cluster kernel in place
    (p3tok_fps); any other D <= 16 runs the general-D kernel (p3tok_fps_nd), which adds the D squares in the order
    torch's CPU sum adds them, so the picks are the reference's for every D."""
    D = int(points.shape[-1])
    if D > 16:
        raise RuntimeError(f"p3tok farthest_point_sampling: D = {D} > 16 coordinates are not supported")
    start = _start(points, start_idx, device_draw=True)
    n = min(int(n_samples), int(points.shape[1]))
    if D != 3:
        return ops.fps_nd(points, start, n)
    return ops.fps(points, start, n)


def square_distance(src: torch.Tensor, dst: torch.Tensor) -> torch.Tensor:
    """sampler.py:47-62 `_square_distance`: (B,N,3), (B,M,3) -> (B,N,M), bit-exact restatement of the reference's
    -2*matmul + |src|^2 + |dst|^2 (may be slightly negative).  knn_point never materialises it; this is for callers
    that use the matrix itself."""
    return ops.square_distance(src, dst)


_square_distance = square_distance     # the reference's (private) name


def index_points(points: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
    """sampler.py:77-94: points (B,N,C), idx (B,S[,k]) -> (B,S[,k],C)."""
    return ops.gather_points(points, idx)


def fps(data: torch.Tensor, number: int, start_idx: Optional[torch.Tensor] = None) -> torch.Tensor:
    """sampler.py:33-45: FPS on xyz then gather all C channels -> (B,number,C).  The second caller
    of the kernel: ScanObjectNN.__init__ downsampling (src/data/scanobjectnn.py:92-97)."""
    idx = ops.fps(data, _start(data, start_idx), int(number))
    return ops.gather_points(data, idx)


def knn_point(nsample: int, xyz: torch.Tensor, new_xyz: torch.Tensor) -> torch.Tensor:
    """sampler.py:64-75: (B,S,nsample) int64.  The reference order is topk(sorted=False)-arbitrary;
    this returns the canonical ascending (distance, index) instance."""
    return ops.knn(xyz, new_xyz, int(nsample), _lib.KNN_APF_SQ, False, False)[0]


def knn_query(pnts: torch.Tensor, cntrds: torch.Tensor, k: int) -> torch.Tensor:
    """The inner `knn` of group_knn (pix4point.py:79-89): cdist + sorted topk -> int32 (B,G,k)."""
    return ops.knn(pnts, cntrds, int(k), _lib.KNN_P4P_CDIST, True, False)[0]


def group_knn(pnts: torch.Tensor, cntrds: torch.Tensor, feats: torch.Tensor, k: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """pix4point.py:56-102: ((B,G,k,3), (B,G,k,D)) - absolute coordinates, no centre subtraction."""
    idx = knn_query(pnts, cntrds, k)
    return ops.group_gather(pnts, feats, idx)


def morton_order(center: torch.Tensor) -> torch.Tensor:
    """MortonEncoder.points_to_morton (apf_utils.py:66-104): (B,G,3) -> (B,G) sorting indices."""
    return ops.morton_order(center)[0]

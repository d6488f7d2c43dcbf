"""GPU parity of the patch embedding at input widths beside BASELINE's: P3Embed with in_channels != 3 (PointViT's
`forward(p, x)` takes any (B, C, N) feature tensor, pix4point.py:234-243 - e.g. xyz + height, xyz + normals) and the APF
PointNet on clouds with more than xyz + height.  Every first-layer route of csrc/embed_tc.cu is reached: the narrow
CUDA-core kernel (cin <= 8), the generic one (cin <= 16) and the gathered tensor-core rows (wider)."""
import numpy as np
import pytest
import torch

from helpers import assert_tokens_close, dev, to_dev
from oracle import oracle
from p3tok import synth
from p3tok.modules import P3Embed, PointNet

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("prec,rtol", [("fp32", 1e-4), ("bf16", 1e-2)])
@pytest.mark.parametrize("D", [1, 4, 5, 6, 13, 14])
def test_p3embed_feature_channels(D, prec, rtol):
    """Stage-0 rows are [xyz | D features] = 4, 7, 8, 9, 16, 17 input channels; two stages at BASELINE's widths 128 / 256."""
    B, N, k = 2, 512, 16
    p = synth.make_cloud("clustered", B, N, 60 + D, 3)
    f = synth.make_points_nd(B, N, D, 160 + D)                 # channel-last features
    sd = synth.p3embed_state(D, 1 / 16, 4, 4, 256, 60 + D)
    mod = P3Embed(in_channels=D, sample_ratio=1 / 16, k=k, embed_dim=256, precision=prec).eval().to(dev())
    mod.load_state_dict(synth.to_torch_state(sd), strict=True)
    starts = [synth.start_indices(B, N, 60 + D, 0), synth.start_indices(B, N // 4, 60 + D, 1)]
    ps, fs = mod(to_dev(p), to_dev(f).transpose(1, 2).contiguous(), [to_dev(s) for s in starts])
    op, of, _ = oracle.p3embed(sd, p, f, starts, k, 2)
    assert mod.channel_list == [D, 128, 256] and mod.out_channels == 256
    for s in (1, 2):
        assert np.array_equal(ps[s].cpu().numpy(), op[s])
        assert tuple(fs[s].shape) == (B, 128 * s, N // 4 ** s)
        assert_tokens_close(fs[s].transpose(1, 2).cpu().numpy(), of[s], rtol * s, f"P3Embed in_channels={D} {prec} stage {s - 1}")


@pytest.mark.parametrize("prec,rtol", [("fp32", 1e-4), ("bf16", 1e-2)])
@pytest.mark.parametrize("C", [5, 6, 9])
def test_apf_pointnet_wide_clouds(C, prec, rtol):
    """AdaptPointFormer(in_channels=C) on clouds that carry more than xyz + height (apf.py:273-310 doubles C for the
    [neighbour - centre | centre] rows: 10, 12, 18 input channels): all C channels are centre-subtracted like the
    reference's Group.forward (apf.py:83-84)."""
    B, N, G, k, E = 2, 512, 24, 16, 128
    x = np.concatenate([synth.make_cloud("uniform", B, N, 70 + C, 3), synth.make_points_nd(B, N, C - 3, 170 + C)], -1)
    st = synth.start_indices(B, N, 70 + C)
    sd = synth.apf_encoder_state(E, 2 * C, 70 + C)
    net = PointNet(E, G, k, 2 * C, precision=prec).eval().to(dev())
    net.encoder.load_state_dict(synth.to_torch_state(sd))
    tok = net(to_dev(x), to_dev(st)).cpu().numpy()
    otok, grp = oracle.pointnet_apf(sd, x, st, G, k)
    assert_tokens_close(tok, otok, rtol, f"APF C={C} {prec}")


@pytest.mark.parametrize("prec,rtol", [("fp32", 1e-4), ("bf16", 1e-2)])
@pytest.mark.parametrize("k", [1, 5, 24, 48, 128])
def test_group_sizes_beside_the_baseline(k, prec, rtol):
    """Neighbourhood sizes that are not 8 / 16 / 32 / 64: one neighbour, sizes that do not divide the 32-row pooling blocks
    or the 256-row tiles, and the kNN kernels' maximum (128) - APF at E = 384 (the fused pair kernels' widths) and a
    two-stage P3Embed at 128 / 256."""
    B, N, G, E = 2, 512, 20, 384
    x = synth.make_cloud("clustered", B, N, 90 + k, 3)
    st = synth.start_indices(B, N, 90 + k)
    sd = synth.apf_encoder_state(E, 6, 90 + k)
    net = PointNet(E, G, k, 6, precision=prec).eval().to(dev())
    net.encoder.load_state_dict(synth.to_torch_state(sd))
    tok = net(to_dev(x), to_dev(st)).cpu().numpy()
    otok, _ = oracle.pointnet_apf(sd, x, st, G, k)
    assert_tokens_close(tok, otok, rtol, f"APF k={k} {prec}")

    sd2 = synth.p3embed_state(3, 1 / 16, 4, 4, 256, 190 + k)
    mod = P3Embed(sample_ratio=1 / 16, k=k, embed_dim=256, precision=prec).eval().to(dev())
    mod.load_state_dict(synth.to_torch_state(sd2), strict=True)
    starts = [synth.start_indices(B, N, 190 + k, 0), synth.start_indices(B, N // 4, 190 + k, 1)]
    ps, fs = mod(to_dev(x), to_dev(x).transpose(1, 2).contiguous(), [to_dev(s) for s in starts])
    op, of, _ = oracle.p3embed(sd2, x, x.copy(), starts, k, 2)
    for s in (1, 2):
        assert np.array_equal(ps[s].cpu().numpy(), op[s])
        assert_tokens_close(fs[s].transpose(1, 2).cpu().numpy(), of[s], rtol * s, f"P3Embed k={k} {prec} stage {s - 1}")

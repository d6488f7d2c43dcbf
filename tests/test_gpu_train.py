"""GPU parity of the training path (SURVEY.md 8f "next" #4): train-mode BatchNorm (batch statistics, running-estimate update)
and the backward through both max-pools, the concat and the gather, through the module drop-ins in .train() - against the
float64 oracle (oracle/train.py) and the fixtures the reference's own forward + autograd produced (tests/golden/apf_train.npz,
p4p_train.npz).  fp32 path: 1e-4 of the largest magnitude of each quantity."""
import os

import numpy as np
import pytest
import torch

import cases
import make_golden
from helpers import dev, to_dev
from oracle import oracle, train as otrain
from p3tok import synth
from p3tok.modules import Encoder, P3Embed, PointNet

pytestmark = pytest.mark.gpu
TOL = 1e-4


def _close(got, ref, what, tol=TOL, scale=None):
    """max|got - ref| <= tol * scale; scale = the largest magnitude of ref, or (parameter gradients) of the largest weight
    gradient - a conv bias in front of a BatchNorm has an exactly-zero gradient, its fp32 noise has no scale of its own."""
    got, ref = np.asarray(got, np.float64), np.asarray(ref, np.float64)
    assert got.shape == ref.shape, (what, got.shape, ref.shape)
    err = np.abs(got - ref).max() / max(np.abs(ref).max() if scale is None else scale, 1e-30)
    print(f"[train parity] {what}: max|err|/max|ref| {err:.2e}")
    assert err <= tol, (what, err)


def _golden_close(got, g, key, what, scale=None):
    """Large matrices are stored in the fixture as row sums and column sums."""
    got = np.asarray(got, np.float64)
    if key in g.files:
        _close(got.reshape(g[key].shape), g[key], what + " vs reference", TOL, scale)
    else:
        m = got.reshape(got.shape[0], -1)
        _close(m.sum(1), g[key + "#rowsum"], what + " row sums vs reference", 2e-4, None if scale is None else scale * m.shape[1] ** 0.5)
        _close(m.sum(0), g[key + "#colsum"], what + " column sums vs reference", 2e-4, None if scale is None else scale * m.shape[0] ** 0.5)


def test_apf_encoder_train_matches_oracle_and_reference(golden_dir):
    c = cases.TRAIN_CASES["apf_train"]
    g = np.load(os.path.join(golden_dir, "apf_train.npz"))
    neigh, gt, sd = make_golden.train_inputs(c)
    enc = Encoder(c["E"], 2 * c["C"]).to(dev()).train()
    enc.load_state_dict(synth.to_torch_state(sd))
    x = to_dev(neigh).requires_grad_(True)
    tok = enc(x)
    (tok * to_dev(gt)).sum().backward()
    otok, ograds, orun = otrain.apf_encoder_train(sd, neigh, gt)
    _close(tok.detach().cpu().numpy(), otok, "tokens vs oracle")
    _close(tok.detach().cpu().numpy(), g["tokens"], "tokens vs reference")
    _close(x.grad.cpu().numpy(), ograds["input"], "grad input vs oracle")
    _golden_close(x.grad.cpu().numpy(), g, "grad.input", "grad input")
    scale = max(np.abs(v).max() for k_, v in ograds.items() if k_.endswith("weight") and "conv" in k_)
    for name, p_ in enc.named_parameters():
        got = p_.grad.cpu().numpy().astype(np.float64)
        ref = ograds[name].reshape(got.shape)
        err = np.abs(got - ref).max() / scale
        print(f"[train parity] grad {name}: {err:.2e} of the largest weight gradient")
        assert err <= TOL, (name, err)
        _golden_close(got, g, "grad." + name, "grad " + name, scale)
    for name, b in enc.named_buffers():
        if "num_batches" in name:
            assert int(b) == 1
            continue
        _close(b.cpu().numpy(), orun[name], "running " + name, 1e-5)
        _close(b.cpu().numpy(), g["running." + name], "running " + name + " vs reference", 1e-5)


def test_p3embed_train_matches_oracle_and_reference(golden_dir):
    c = cases.P4P_TRAIN_CASES["p4p_train"]
    g = np.load(os.path.join(golden_dir, "p4p_train.npz"))
    x, start, sd, go = make_golden.p4p_train_inputs(c)
    m = P3Embed(in_channels=3, sample_ratio=0.25, scale=4, k=c["k"], layers=4, embed_dim=c["W"]).to(dev()).train()
    m.load_state_dict(synth.to_torch_state(sd))
    p = to_dev(x).requires_grad_(True)
    fe = to_dev(np.ascontiguousarray(x.transpose(0, 2, 1))).requires_grad_(True)
    _, fs = m(p, fe, [to_dev(start)])
    out = fs[-1].transpose(1, 2)
    (out * to_dev(go)).sum().backward()
    _close(out.detach().cpu().numpy(), g["out"], "stage output vs reference")
    _close(p.grad.cpu().numpy(), g["grad.p"], "grad p vs reference")
    _close(fe.grad.transpose(1, 2).cpu().numpy(), g["grad.f"], "grad f vs reference")
    scale = max(np.abs(g[k_]).max() for k_ in g.files if k_.startswith("grad.convs") and k_.endswith("weight"))
    for name, q in m.named_parameters():
        err = np.abs(q.grad.cpu().numpy().astype(np.float64) - g["grad." + name]).max() / scale
        print(f"[train parity] grad {name}: {err:.2e} of the largest weight gradient")
        assert err <= TOL, (name, err)
    for name, b in m.named_buffers():
        if "num_batches" not in name:
            _close(b.cpu().numpy(), g["running." + name], "running " + name + " vs reference", 1e-5)


def test_pointnet_train_end_to_end_against_oracle():
    """PointNet.train(): FPS / kNN / Morton indices (no gradient), gathered rows, train-mode Encoder - at a shape with
    ragged tiles (rows not a multiple of the 64 / 128-row kernel tiles) and a 4-channel input."""
    B, N, C, G, k, E = 3, 500, 4, 21, 12, 48
    x = synth.make_cloud("clustered", B, N, 91, C)
    st = synth.start_indices(B, N, 91)
    sd = synth.apf_encoder_state(E, 2 * C, 91)
    net = PointNet(E, G, k, 2 * C).to(dev()).train()
    net.encoder.load_state_dict(synth.to_torch_state(sd))
    gt = (synth.uniform01(91, B * G * E, 3).reshape(B, G, E) - 0.5).astype(np.float32)
    tok = net(to_dev(x), to_dev(st))
    (tok * to_dev(gt)).sum().backward()
    grp = oracle.group_apf(x, st, G, k)
    otok, ograds, orun = otrain.apf_encoder_train(sd, grp["neigh"], gt)
    _close(tok.detach().cpu().numpy(), otok, "PointNet.train tokens vs oracle")
    scale = max(np.abs(v).max() for k_, v in ograds.items() if k_.endswith("weight") and "conv" in k_)
    for name, p_ in net.encoder.named_parameters():
        got = p_.grad.cpu().numpy().astype(np.float64)
        err = np.abs(got - ograds[name].reshape(got.shape)).max() / scale
        assert err <= TOL, (name, err)
    # a second step keeps accumulating the running estimates like nn.BatchNorm (momentum 0.1)
    rm0 = net.encoder.first_conv[1].running_mean.clone()
    net(to_dev(x), to_dev(st))
    assert int(net.encoder.first_conv[1].num_batches_tracked) == 2 and not torch.equal(rm0, net.encoder.first_conv[1].running_mean)
    # eval() afterwards uses the updated running statistics (folded weights are rebuilt)
    net.eval()
    with torch.no_grad():
        te = net(to_dev(x), to_dev(st))
    sd2 = {k_: v.detach().cpu().numpy() for k_, v in net.encoder.state_dict().items()}
    etok, _ = oracle.pointnet_apf(sd2, x, st, G, k)
    _close(te.cpu().numpy(), etok, "eval after train vs oracle", 1e-4)


def test_train_kernels_building_blocks():
    """linear_tn / colstats / group max with arg-max / scatter against torch on ragged shapes."""
    from p3tok import train as T
    torch.manual_seed(0)
    for (M, N, K) in ((1, 3, 2), (37, 5, 6), (1000, 70, 130), (4099, 256, 64)):
        dy, x = torch.randn(M, N, device=dev()), torch.randn(M, K, device=dev())
        ref = dy.double().t() @ x.double()
        assert float((T.linear_tn(dy, x).double() - ref).abs().max()) <= 1e-5 * max(float(ref.abs().max()), 1.0)
        s, q = T.colstats(dy)
        assert torch.allclose(s, dy.double().sum(0), atol=1e-4) and torch.allclose(q, (dy.double() ** 2).sum(0), rtol=1e-5, atol=1e-4)
    v = torch.randn(7 * 9, 33, device=dev())
    v[3, 5] = v[4, 5] = 100.0                               # a tie: the FIRST maximum wins like torch.max
    m, a = T.group_max_arg(v, 9)
    rm, ra = v.view(7, 9, 33).max(1)
    assert torch.equal(m, rm) and torch.equal(a.long(), ra)
    d = torch.randn(7, 33, device=dev())
    dx = T.group_max_bwd(d, a, 9)
    ref = torch.zeros(7, 9, 33, device=dev()).scatter_(1, ra.unsqueeze(1), d.unsqueeze(1)).view(63, 33)
    assert torch.equal(dx, ref)
    assert torch.allclose(T.group_sum(v, 9), v.view(7, 9, 33).sum(1), atol=1e-4)

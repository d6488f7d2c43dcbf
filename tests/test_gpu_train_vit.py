"""GPU parity of the training path through the token consumer (SURVEY.md 8f "next" #4 on top of #3): the APFViTLayer stack,
encoder_norm, the max over tokens, dropout and the ClassificationHead under autograd (csrc/train_vit.cu, p3tok/train_vit.py)
against the float64 oracle (oracle/train.py) and the fixtures the reference's own modules + autograd produced
(tests/golden/vit_train.npz, vit_train_full.npz, vit_train_masked.npz).  fp32 path: 1e-4 of the largest magnitude of each
quantity; building blocks also against the plain PyTorch fp32 evaluation of the same op on the device."""
import os

import numpy as np
import pytest
import torch

import cases
import make_golden
from helpers import dev, to_dev
from oracle import train as otrain
from p3tok import synth, train_vit
from p3tok.apf_model import AdaptPointFormer, APFViTLayer, ClassificationHead, run_blocks

pytestmark = pytest.mark.gpu
TOL = 1e-4


def _close(got, ref, what, tol=TOL, scale=None):
    got, ref = np.asarray(got, np.float64), np.asarray(ref, np.float64)
    assert got.shape == ref.shape, (what, got.shape, ref.shape)
    assert np.isfinite(got).all(), what
    err = np.abs(got - ref).max() / max(np.abs(ref).max() if scale is None else scale, 1e-30)
    print(f"[train parity] {what}: max|err|/scale {err:.2e}")
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


def _stack(sd, D, heads, depth, train=False, p_adapter=0.0, dpr=0.0):
    tsd = synth.to_torch_state(sd)
    layers = []
    for i in range(depth):
        l = APFViTLayer(D, heads, drop_path=dpr, dropout=p_adapter)
        l.load_state_dict({k[len(f"blocks.{i}."):]: v for k, v in tsd.items() if k.startswith(f"blocks.{i}.")}, strict=True)
        layers.append(l.to(dev()).train(train))
    norm = torch.nn.LayerNorm(D)
    norm.load_state_dict({"weight": tsd["encoder_norm.weight"], "bias": tsd["encoder_norm.bias"]})
    return layers, norm.to(dev()).train(train)


def test_gradient_through_frozen_blocks_matches_oracle_and_reference(golden_dir):
    """What the reference's training step needs from the frozen layers (apf.py:335-346): eval-mode blocks, frozen parameters, a
    token tensor that requires grad -> d tokens and the encoder_norm gradients."""
    c = cases.VIT_TRAIN_CASES["vit_train"]
    g = np.load(os.path.join(golden_dir, "vit_train.npz"))
    sd = synth.apf_vit_state(c["D"], c["depth"], 15, c["seed"])
    tok = synth.vit_tokens(c["B"], c["G"], c["D"], c["seed"])
    gp = (synth.uniform01(c["seed"], c["B"] * c["D"], 35).reshape(c["B"], c["D"]) - 0.5).astype(np.float32)
    layers, norm = _stack(sd, c["D"], c["heads"], c["depth"], train=False)
    for l in layers:
        for p_ in l.parameters():
            p_.requires_grad_(False)
    x = to_dev(tok).requires_grad_(True)
    _, pooled = run_blocks(layers, x, norm)
    (pooled * to_dev(gp)).sum().backward()
    po, dx, gn = otrain.apf_vit_backward(sd, tok, c["depth"], c["heads"], gp)
    _close(pooled.detach().cpu().numpy(), po, "pooled vs oracle")
    _close(pooled.detach().cpu().numpy(), g["pooled"], "pooled vs reference")
    _close(x.grad.cpu().numpy(), dx, "d tokens vs oracle")
    _close(x.grad.cpu().numpy(), g["grad_tokens"], "d tokens vs reference")
    _close(norm.weight.grad.cpu().numpy(), g["grad_norm_w"], "encoder_norm.weight grad vs reference")
    _close(norm.bias.grad.cpu().numpy(), g["grad_norm_b"], "encoder_norm.bias grad vs reference")
    assert all(p_.grad is None for l in layers for p_ in l.parameters())
    # without a gradient request the same modules take the serving (tensor-core) path: bf16-level agreement only
    with torch.no_grad():
        _, served = run_blocks(layers, to_dev(tok), norm)
    _close(served.cpu().numpy(), po, "serving path vs oracle", 2e-2)


@pytest.mark.parametrize("name", list(cases.VIT_FULL_TRAIN_CASES))
def test_full_consumer_train_step_matches_oracle_and_reference(golden_dir, name):
    """Blocks (every parameter asks for a gradient) -> encoder_norm -> max -> dropout -> ClassificationHead in train mode, with
    the forced keep masks of the fixture."""
    c = cases.VIT_FULL_TRAIN_CASES[name]
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    sd, tok, gl, masks = make_golden.vit_full_train_inputs(c)
    layers, norm = _stack(sd, c["D"], c["heads"], c["depth"], train=True, p_adapter=c["p_adapter"])
    head = ClassificationHead(c["D"], c["classes"])
    head.load_state_dict({k[len("head."):]: v for k, v in synth.to_torch_state(sd).items() if k.startswith("head.")}, strict=True)
    head = head.to(dev()).train()
    md = lambda m: None if m is None else to_dev(m)
    lm = [(None, md(masks["adapter"][i]), None) for i in range(c["depth"])]
    if all(m[1] is None for m in lm):
        lm = None
    x = to_dev(tok).requires_grad_(True)
    _, pooled = train_vit.blocks_train(layers, x, norm, masks=lm)
    hin = train_vit.dropout(pooled, 0.1, True, mask=md(masks["pool"])) if masks["pool"] is not None else pooled
    hm = (md(masks["head"][0]), md(masks["head"][1])) if masks["head"][0] is not None else None
    logits = train_vit.head_train(head, hin, masks=hm) if hm is not None else _head_no_dropout(head, hin)
    (logits * to_dev(gl)).sum().backward()
    _close(pooled.detach().cpu().numpy(), g["pooled"], "pooled vs reference")
    _close(logits.detach().cpu().numpy(), g["logits"], "logits vs reference")
    _close(x.grad.cpu().numpy(), g["grad.tokens"], "d tokens vs reference")
    scale = max(np.abs(g[k_]).max() for k_ in g.files if k_.startswith("grad.") and k_.endswith("weight") and g[k_].ndim == 2)
    for i, l in enumerate(layers):
        for n, p_ in l.named_parameters():
            assert p_.grad is not None, n
            _golden_close(p_.grad.cpu().numpy(), g, f"grad.blocks.{i}.{n}", f"grad blocks.{i}.{n}", scale)
    for n, p_ in norm.named_parameters():
        _golden_close(p_.grad.cpu().numpy(), g, "grad.encoder_norm." + n, "grad encoder_norm." + n, scale)
    for n, p_ in head.named_parameters():
        _golden_close(p_.grad.cpu().numpy(), g, "grad.head." + n, "grad head." + n, scale)
    for n, b in head.named_buffers():
        if "num_batches" in n:
            assert int(b) == 1
        else:
            _close(b.cpu().numpy(), g["running.head." + n], "running head." + n + " vs reference", 1e-5)


def _head_no_dropout(head, x):
    """The fixture without dropout: the head's nn.Dropout rates are fixed at 0.4 in the reference, so 'off' = all-ones masks."""
    m = head.mlp_head
    ones = (torch.ones((x.shape[0], m[0].out_features), device=x.device), torch.ones((x.shape[0], m[4].out_features), device=x.device))
    return train_vit.head_train(head, x, masks=ones)


def test_drop_path_and_dropout_masks_against_oracle():
    """All three stochastic regularisers of a layer as explicit keep masks (timm DropPath: one draw per cloud and branch, scaled
    by 1 / keep) at a ragged shape; every parameter gradient against the float64 oracle."""
    B, G, D, heads, depth = 5, 37, 96, 3, 2
    sd = synth.apf_vit_state(D, depth, 15, 211)
    tok = synth.vit_tokens(B, G, D, 211)
    gp = (synth.uniform01(211, B * D, 5).reshape(B, D) - 0.5).astype(np.float32)
    rs = np.random.RandomState(5)
    keep = lambda shape, p: ((rs.rand(*shape) >= p) / (1.0 - p)).astype(np.float32)
    masks = [(keep((B,), 0.3), keep((B * G, 64), 0.2), keep((B,), 0.3)) for _ in range(depth)]
    layers, norm = _stack(sd, D, heads, depth, train=True, p_adapter=0.2, dpr=0.3)
    x = to_dev(tok).requires_grad_(True)
    y, pooled = train_vit.blocks_train(layers, x, norm, masks=[tuple(to_dev(m) for m in t) for t in masks])
    (pooled * to_dev(gp)).sum().backward()
    po, dx, og = otrain.apf_vit_backward(sd, tok, depth, heads, gp, masks, param_grads=True)
    _close(pooled.detach().cpu().numpy(), po, "pooled (masked) vs oracle")
    _close(x.grad.cpu().numpy(), dx, "d tokens (masked) vs oracle")
    scale = max(np.abs(v).max() for k_, v in og.items() if k_.endswith("weight") and v.ndim == 2)
    for i, l in enumerate(layers):
        for n, p_ in l.named_parameters():
            _close(p_.grad.cpu().numpy().reshape(-1), og[f"blocks.{i}.{n}"].reshape(-1), f"grad blocks.{i}.{n} vs oracle", TOL, scale)
    for n, p_ in norm.named_parameters():
        _close(p_.grad.cpu().numpy(), og["encoder_norm." + n], "grad encoder_norm." + n + " vs oracle", TOL, scale)
    # drawn masks: a train-mode call without explicit masks draws its own (different outputs call to call, finite gradients)
    y1 = train_vit.blocks_train(layers, to_dev(tok), None)[0]
    y2 = train_vit.blocks_train(layers, to_dev(tok), None)[0]
    assert torch.isfinite(y1).all() and not torch.equal(y1, y2)


@pytest.mark.parametrize("B,G,heads,hd", [(2, 128, 12, 32), (1, 197, 3, 64), (3, 1, 2, 32), (2, 33, 1, 16), (1, 257, 2, 64)])
def test_attention_kernels_match_torch(B, G, heads, hd):
    D = heads * hd
    gen = torch.Generator(device="cpu").manual_seed(B * 1000 + G)
    qkv = torch.randn(B * G, 3 * D, generator=gen).to(dev())
    do = torch.randn(B * G, D, generator=gen).to(dev())
    o, P = train_vit.attn_fwd(qkv, B, G, heads)
    dqkv = train_vit.attn_bwd(qkv, P, do, B, G, heads)
    t = qkv.double().clone().requires_grad_(True)
    q, k, v = t.reshape(B, G, 3, heads, hd).permute(2, 0, 3, 1, 4)
    ref = ((q @ k.transpose(-2, -1)) * hd ** -0.5).softmax(-1) @ v
    ref = ref.transpose(1, 2).reshape(B * G, D)
    (ref * do.double()).sum().backward()
    _close(o.cpu().numpy(), ref.detach().cpu().numpy(), f"attention forward G={G} hd={hd}", 2e-5)
    _close(dqkv.cpu().numpy(), t.grad.cpu().numpy(), f"attention backward G={G} hd={hd}", 2e-5)


def test_layernorm_and_elementwise_kernels_match_torch():
    gen = torch.Generator(device="cpu").manual_seed(3)
    for M, D in ((1000, 384), (7, 33), (1, 768), (4097, 64)):
        x = (torch.randn(M, D, generator=gen) * 2 + 0.5).to(dev())
        w, b = (torch.rand(D, generator=gen) + 0.5).to(dev()), torch.randn(D, generator=gen).to(dev())
        dy = torch.randn(M, D, generator=gen).to(dev())
        y, mean, rstd = train_vit.ln_fwd(x, w, b, 1e-5)
        xt = x.double().requires_grad_(True)
        wt, bt = w.double().requires_grad_(True), b.double().requires_grad_(True)
        ref = torch.nn.functional.layer_norm(xt, (D,), wt, bt, 1e-5)
        (ref * dy.double()).sum().backward()
        _close(y.cpu().numpy(), ref.detach().cpu().numpy(), f"LayerNorm forward {M}x{D}", 1e-5)
        base = torch.ones_like(x)
        dx = train_vit.ln_bwd(dy, x, mean, rstd, w, into=base.clone())
        _close((dx - base).cpu().numpy(), xt.grad.cpu().numpy(), f"LayerNorm backward {M}x{D}", 2e-5)
        gw, gb = train_vit.ln_param_grad(dy, x, mean, rstd)
        _close(gw.cpu().numpy(), wt.grad.cpu().numpy(), f"LayerNorm weight grad {M}x{D}", 2e-5)
        _close(gb.cpu().numpy(), bt.grad.cpu().numpy(), f"LayerNorm bias grad {M}x{D}", 2e-5)
    from p3tok import _lib
    a = torch.randn(5, 1001, generator=gen).to(dev()) * 3
    b = torch.randn(5, 1001, generator=gen).to(dev())
    at = a.double().requires_grad_(True)
    gl = torch.nn.functional.gelu(at)
    (gl * b.double()).sum().backward()
    _close(train_vit.ew(_lib.EW_GELU, a).cpu().numpy(), gl.detach().cpu().numpy(), "gelu", 1e-6)
    _close(train_vit.ew(_lib.EW_GELU_BWD, a, b).cpu().numpy(), at.grad.cpu().numpy(), "gelu backward", 1e-6)
    _close(train_vit.ew(_lib.EW_RELU_BWD, a, b).cpu().numpy(), (b * (a > 0)).cpu().numpy(), "relu backward", 0)
    _close(train_vit.axpby(0.5, a, -2.0, b).cpu().numpy(), (0.5 * a - 2.0 * b).cpu().numpy(), "axpby", 1e-6)
    per = torch.tensor([0.0, 2.0, 1.0, 0.0, 4.0], device=dev())
    _close(train_vit.mask_mul(a, per, 1001).cpu().numpy(), (a * per[:, None]).cpu().numpy(), "per-cloud mask", 0)
    with pytest.raises(RuntimeError):
        train_vit.ln_fwd(torch.zeros(4, 8), None, None, 1e-5)                     # CPU tensor: no fallback


def test_adaptpointformer_training_step():
    """The reference's training configuration end to end on the kernels: AdaptPointFormer.train() with its _freeze() rule
    (apf.py:335-346) - logits, a loss, backward: gradients exactly for point_encoder / encoder_norm / head, BatchNorm buffers
    advance, and with every stochastic rate at 0 the gradient of the tokens agrees with the oracle's block-stack backward."""
    B, N, G, k, E = 4, 256, 16, 8, 384
    m = AdaptPointFormer(num_classes=7, embedding_dim=E, npoint=G, nsample=k, in_channels=3, dropout_rate=0.0, dropout_path_rate=0.0,
                         precision="fp32")
    sd = synth.apf_vit_state(E, 12, 7, 17)
    sd.update({"point_encoder.encoder." + k_: v for k_, v in synth.apf_encoder_state(E, 6, 17).items()})
    m.load_state_dict(synth.to_torch_state(sd), strict=True)
    m = m.to(dev()).train()
    m._freeze()
    for d_ in (m.head.mlp_head[3], m.head.mlp_head[7]):
        d_.p = 0.0
    x = to_dev(synth.make_cloud("uniform", B, N, 17, 3))
    st = to_dev(synth.start_indices(B, N, 17))
    tok = m.point_encoder(x, st)
    tok.retain_grad()
    cache = {}
    pooled = run_blocks(m.blocks, tok, m.encoder_norm, cache)[1]
    logits = m.head(pooled)
    target = torch.arange(B, device=dev()) % 7
    # loss = sum(logits * onehot-ish weights): a linear functional keeps the oracle comparison exact
    gl = torch.nn.functional.one_hot(target, 7).float() - 1.0 / 7
    (logits * gl).sum().backward()
    for n, p_ in m.named_parameters():
        trainable = ("head" in n) or ("encoder" in n)
        assert p_.requires_grad == trainable, n
        assert (p_.grad is not None) == trainable, n
        if trainable:
            assert torch.isfinite(p_.grad).all(), n
    assert int(m.head.mlp_head[1].num_batches_tracked) == 1 and int(m.point_encoder.encoder.first_conv[1].num_batches_tracked) == 1
    # oracle: head backward on the kernel's pooled features, then the block-stack backward on the kernel's tokens
    _, hg, _ = otrain.head_train(sd, pooled.detach().cpu().numpy(), gl.cpu().numpy())
    po, dx, gn = otrain.apf_vit_backward(sd, tok.detach().cpu().numpy(), 12, 12, hg["input"])
    _close(pooled.detach().cpu().numpy(), po, "pooled (12 layers) vs oracle")
    _close(tok.grad.cpu().numpy(), dx, "d tokens through 12 layers vs oracle")
    _close(m.encoder_norm.weight.grad.cpu().numpy(), gn["encoder_norm.weight"], "encoder_norm.weight grad vs oracle")
    # the module's own forward does the same thing (default rates: dropout / DropPath masks drawn), and an optimiser step runs
    m2 = AdaptPointFormer(num_classes=7, embedding_dim=E, npoint=G, nsample=k, in_channels=3, precision="fp32").to(dev()).train()
    m2._freeze()
    opt = torch.optim.SGD([p_ for p_ in m2.parameters() if p_.requires_grad], lr=1e-3)
    out = m2(x, st)
    assert out.shape == (B, 7)
    torch.nn.functional.cross_entropy(out, target).backward()
    opt.step()
    m2.eval()
    with torch.no_grad():
        assert torch.isfinite(m2(x, st)).all()


def test_pix4point_block_loop_train_matches_oracle_and_fixture(golden_dir):
    """PointViT's block loop + final norm + 'max,cls' features under autograd (pix4point.py:254-271): feats, pos and every
    parameter against the float64 oracle and the torch.nn.TransformerEncoderLayer-autograd fixture."""
    from p3tok.p4p_model import TimmBlock
    c = cases.P4P_VIT_TRAIN_CASES["p4p_vit_train"]
    g = np.load(os.path.join(golden_dir, "p4p_vit_train.npz"))
    feats, pos, sd = make_golden.p4p_vit_inputs(c)
    D = c["D"]
    gg = (synth.uniform01(c["seed"], c["B"] * 2 * D, 39).reshape(c["B"], 2 * D) - 0.5).astype(np.float32)
    tsd = synth.to_torch_state(sd)
    blocks = []
    for i in range(c["depth"]):
        b = TimmBlock(D, c["heads"])
        b.load_state_dict({k[len(f"vit.blocks.{i}."):]: v for k, v in tsd.items() if k.startswith(f"vit.blocks.{i}.")}, strict=True)
        blocks.append(b.to(dev()).train())
    norm = torch.nn.LayerNorm(D, eps=1e-6)
    norm.load_state_dict({"weight": tsd["vit.norm.weight"], "bias": tsd["vit.norm.bias"]})
    norm = norm.to(dev()).train()
    x, p = to_dev(feats).requires_grad_(True), to_dev(pos).requires_grad_(True)
    out = train_vit.timm_blocks_train(blocks, norm, x, p)
    glob = torch.cat([train_vit.TokenMaxFn.apply(out, 1), out[:, 0, :]], 1)
    (glob * to_dev(gg)).sum().backward()
    og, dx, dp, grads = otrain.pointvit_backward(sd, feats, pos, c["depth"], c["heads"], gg)
    _close(glob.detach().cpu().numpy(), og, "global features vs oracle")
    _close(glob.detach().cpu().numpy(), g["glob"], "global features vs fixture")
    _close(x.grad.cpu().numpy(), dx, "d feats vs oracle")
    _close(p.grad.cpu().numpy(), dp, "d pos vs oracle")
    _golden_close(x.grad.cpu().numpy(), g, "grad.feats", "d feats")
    _golden_close(p.grad.cpu().numpy(), g, "grad.pos", "d pos")
    scale = max(np.abs(v).max() for k_, v in grads.items() if k_.endswith("weight") and v.ndim == 2)
    for i, b in enumerate(blocks):
        for n, q in b.named_parameters():
            _close(q.grad.cpu().numpy().reshape(-1), grads[f"vit.blocks.{i}.{n}"].reshape(-1), f"grad vit.blocks.{i}.{n} vs oracle", TOL, scale)
            _golden_close(q.grad.cpu().numpy(), g, f"grad.vit.blocks.{i}.{n}", f"grad vit.blocks.{i}.{n}", scale)
    for n, q in norm.named_parameters():
        _close(q.grad.cpu().numpy(), grads["vit.norm." + n], "grad vit.norm." + n + " vs oracle", TOL, scale)


def test_pointvit_training_step():
    """PointViT.train(): P3Embed (batch-statistics BatchNorm) -> proj / pos_embed / cls concat -> block loop -> global features,
    loss, backward: every parameter the reference trains receives a finite gradient; frozen=True leaves `vit.*` without one
    (pix4point.py:229-233); the token head's gradients agree with the plain PyTorch evaluation of the same layers."""
    from p3tok.p4p_model import PointViT
    B, N = 3, 256
    x = to_dev(synth.make_cloud("uniform", B, N, 23, 3))
    st = [to_dev(synth.start_indices(B, N, 23))]
    for frozen in (False, True):
        torch.manual_seed(1)
        m = PointViT(embed_dim=64, depth=2, num_heads=2, k_neighbors=8, frozen=frozen, precision="fp32").to(dev()).train()
        with torch.no_grad():
            for q in (m.cls_token, m.cls_pos):
                q.normal_(0, 0.02)
        gf = m.forward_cls_feat(x, None, st)
        assert gf.shape == (B, 128)
        (gf * torch.linspace(-1, 1, 128, device=dev())).sum().backward()
        for n, q in m.named_parameters():
            if n == "vit.pos_embed":
                continue                                     # timm's own position table: PointViT reads only its first row, at construction
            want = q.requires_grad
            assert want == (not (frozen and "vit" in n)), n
            assert (q.grad is not None) == want, n
            if want:
                assert torch.isfinite(q.grad).all() and float(q.grad.abs().max()) > 0, n
        assert int(m.patch_embed.convs[0][0][2].num_batches_tracked) == 1
    # token head against torch on the device
    tokm = m.__dict__["_tok"]
    tk = torch.randn(B, 16, m.patch_embed.out_channels, device=dev(), requires_grad=True)
    ctr = torch.rand(B, 16, 3, device=dev())
    for q in m.parameters():
        q.requires_grad_(True)
        q.grad = None
    feats, pos = train_vit.token_head_train(tokm, tk, ctr)
    w1 = torch.randn_like(feats)
    ((feats + 0.5 * pos) * w1).sum().backward()
    got = {n: q.grad.clone() for n, q in (("proj.weight", m.proj.weight), ("pos0.weight", m.pos_embed[0].weight), ("pos2.bias", m.pos_embed[2].bias),
                                          ("cls_token", m.cls_token), ("cls_pos", m.cls_pos), ("tokens", tk))}
    for q in list(m.parameters()) + [tk]:
        q.grad = None
    F = torch.nn.functional
    xr = F.linear(tk, m.proj.weight, m.proj.bias)
    pr = F.linear(F.gelu(F.linear(ctr, m.pos_embed[0].weight, m.pos_embed[0].bias)), m.pos_embed[2].weight, m.pos_embed[2].bias)
    fr = torch.cat([m.cls_token.expand(B, -1, -1), xr], 1)
    prr = torch.cat([m.cls_pos.expand(B, -1, -1), pr], 1)
    ((fr + 0.5 * prr) * w1).sum().backward()
    _close(feats.detach().cpu().numpy(), fr.detach().cpu().numpy(), "token head feats vs torch", 1e-5)
    _close(pos.detach().cpu().numpy(), prr.detach().cpu().numpy(), "token head pos vs torch", 1e-5)
    ref = {"proj.weight": m.proj.weight.grad, "pos0.weight": m.pos_embed[0].weight.grad, "pos2.bias": m.pos_embed[2].bias.grad,
           "cls_token": m.cls_token.grad, "cls_pos": m.cls_pos.grad, "tokens": tk.grad}
    for n in got:
        _close(got[n].cpu().numpy(), ref[n].cpu().numpy(), "token head grad " + n + " vs torch", 2e-5)


def test_clshead_and_pix4point_model():
    """ClsHead (pix4point.py:295-325) eval / train against the plain PyTorch evaluation of its own nn.Sequential container, and
    a full Pix4Point training step (pix4point.py:328-437) on the kernels."""
    import copy
    from p3tok.p4p_model import ClsHead, Pix4Point
    torch.manual_seed(7)
    head = ClsHead(in_channels=96, num_classes=11, mlps=[64, 48, 32], dropout=0.0).to(dev())
    with torch.no_grad():
        for mod in head.head:
            if isinstance(mod, torch.nn.BatchNorm1d):
                mod.running_mean.normal_(0, 0.2); mod.running_var.uniform_(0.5, 1.5); mod.weight.uniform_(0.5, 1.5); mod.bias.normal_(0, 0.1)
    ref = copy.deepcopy(head.head)
    x = torch.randn(10, 96, device=dev())
    _close(head.eval()(x).cpu().numpy(), ref.eval()(x).detach().cpu().numpy(), "ClsHead eval vs torch", 1e-5)
    xa, xb = x.clone().requires_grad_(True), x.clone().requires_grad_(True)
    w = torch.randn(10, 11, device=dev())
    (head.train()(xa) * w).sum().backward()
    (ref.train()(xb) * w).sum().backward()
    _close(xa.grad.cpu().numpy(), xb.grad.cpu().numpy(), "ClsHead train d input vs torch")
    scale = max(float(q.grad.abs().max()) for q in ref.parameters() if q.ndim == 2)
    for (n, q), (_, r) in zip(head.head.named_parameters(), ref.named_parameters()):
        _close(q.grad.cpu().numpy(), r.grad.cpu().numpy(), "ClsHead train grad " + n + " vs torch", TOL, scale)
    for (n, q), (_, r) in zip(head.head.named_buffers(), ref.named_buffers()):
        _close(q.float().cpu().numpy(), r.float().cpu().numpy(), "ClsHead running " + n + " vs torch", 1e-5)
    # the whole model: a training step, then serving
    B, N = 4, 256
    pts = to_dev(synth.make_cloud("clustered", B, N, 29, 3))
    st = [to_dev(synth.start_indices(B, N, 29))]
    m = Pix4Point(num_classes=5, embed_dim=64, depth=2, num_heads=2, k_neighbors=8, precision="fp32").to(dev()).train()
    opt = torch.optim.AdamW(m.get_param_groups(), lr=1e-3, weight_decay=0.05)
    before = m.cls_head.head[0].weight.detach().clone()
    logits = m(pts, st)
    assert logits.shape == (B, 5)
    torch.nn.functional.cross_entropy(logits, torch.arange(B, device=dev()) % 5).backward()
    opt.step()
    assert not torch.equal(before, m.cls_head.head[0].weight)
    assert all(torch.isfinite(q.grad).all() for q in m.parameters() if q.grad is not None)
    with torch.no_grad():
        assert torch.isfinite(m.eval()(pts, st)).all()


def test_tensor_core_training_gemms():
    """P3TOK_TRAIN_TC / set_tensor_core_gemms(1): the forward and dX products of the training path through the bf16x3 tensor-core
    GEMM (p3tok_linear_x3_f32).  The product itself against float64, then a train-mode Encoder step against the oracle (tokens
    1e-4 of max; gradients in aggregate, see below)."""
    from p3tok import train
    from p3tok.modules import Encoder
    gen = torch.Generator(device="cpu").manual_seed(11)
    for M, K, N in ((4096, 384, 1152), (1500, 6, 256), (2048, 1600, 384), (1024, 131, 64)):
        a = torch.randn(M, K, generator=gen).to(dev())
        w = (torch.randn(N, K, generator=gen) / K ** 0.5).to(dev())
        b = torch.randn(N, generator=gen).to(dev())
        prev = train.set_tensor_core_gemms(1)
        try:
            got = train.linear(a, w, b)
        finally:
            train.set_tensor_core_gemms(prev)
        ref = a.double() @ w.double().t() + b.double()
        _close(got.cpu().numpy(), ref.cpu().numpy(), f"bf16x3 linear {M}x{K}x{N} vs float64", 3e-5)
    # 2048 rows (>= the 1024-row threshold below which the SGEMM is kept), ragged against the 256-row tiles of the GEMM
    from oracle import oracle
    B, N, C, G, k, E = 3, 512, 3, 43, 16, 64
    xs = synth.make_cloud("uniform", B, N, 77, C)
    neigh = oracle.group_apf(xs, synth.start_indices(B, N, 77), G, k)["neigh"].astype(np.float32)
    sd = synth.apf_encoder_state(E, 2 * C, 77)
    gt = (synth.uniform01(77, B * G * E, 31).reshape(B, G, E) - 0.5).astype(np.float32)
    enc = Encoder(E, 2 * C).to(dev()).train()
    enc.load_state_dict(synth.to_torch_state(sd))
    x = to_dev(neigh).requires_grad_(True)
    from p3tok import ops
    prev = train.set_tensor_core_gemms(1)
    n0 = ops.kernel_launches()
    try:
        tok = enc(x)
        (tok * to_dev(gt)).sum().backward()
    finally:
        train.set_tensor_core_gemms(prev)
    launches_tc = ops.kernel_launches() - n0
    otok, ograds, _ = otrain.apf_encoder_train(sd, neigh, gt)
    _close(tok.detach().cpu().numpy(), otok, "tensor-core training: tokens vs oracle", 1e-4)
    # Gradients are ROUTED by the two arg-max pools (8256 picks each here): tokens that differ by 1e-5 pick another row of a
    # near-tie in a handful of places, and one moved pick changes a weight gradient by ~1 % in the Frobenius norm (measured:
    # every bf16x3 product of this step is within 7e-6 of float64 on its real operands, tests/_dbg_tc_train.py, while the
    # gradients differ from the fp32-SGEMM run by 1-3 % - and the fp32-SGEMM run itself moves by 0.7-1.6 % when one layer's
    # weights are perturbed by 1e-5: the sensitivity is the function's, not the GEMM's).  So gradients are held to an aggregate
    # bound here; the element-wise 1e-4 contract is the CUDA-core default's.
    def fro(a, b):
        a, b = np.asarray(a, np.float64).reshape(-1), np.asarray(b, np.float64).reshape(-1)
        return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))
    e = fro(x.grad.cpu().numpy(), ograds["input"])
    print(f"[train parity] tensor-core training: grad input vs oracle: relative Frobenius error {e:.2e}")
    assert e <= 6e-2, e
    for name, p_ in enc.named_parameters():
        if name.endswith("weight"):
            e = fro(p_.grad.cpu().numpy(), ograds[name])
            print(f"[train parity] tensor-core training: grad {name} vs oracle: relative Frobenius error {e:.2e}")
            assert e <= 6e-2, (name, e)
    # the tensor-core path really ran: every bf16x3 product adds two split kernels to the launch count of the SGEMM path
    for p_ in enc.parameters():
        p_.grad = None
    n0 = ops.kernel_launches()
    (enc(x.detach().requires_grad_(True)) * to_dev(gt)).sum().backward()
    assert launches_tc > ops.kernel_launches() - n0

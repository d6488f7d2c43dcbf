"""GPU parity of the ViT block stack that consumes the APF tokens (SURVEY.md 8f "next" #3): building blocks against
torch, the stack against the float64 oracle and the reference-generated golden vectors, through the module drop-ins
(p3tok.apf_model) and the C ABI.  Tolerance: the bf16 contract of the north star (rtol 1e-2, helpers.assert_tokens_close),
plus a tighter bound against a torch model of the kernels' own arithmetic (helpers.vit_bf16_emulation)."""
import os

import numpy as np
import pytest
import torch

import cases
from helpers import assert_tokens_close, dev, rel_err, to_dev, vit_bf16_emulation
from oracle import oracle
from p3tok import ops, synth
from p3tok.apf_model import AdaptPointFormer, APFViTLayer, ClassificationHead, fold_vit_layer, run_blocks

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("M,D", [(1, 64), (37, 384), (1000, 768), (5, 1024), (64, 100)])
def test_layernorm_block(M, D):
    torch.manual_seed(M + D)
    x = torch.randn(M, D, device=dev()) * 3 + 1
    w, b = torch.rand(D, device=dev()) + 0.5, torch.randn(D, device=dev()) * 0.1
    got = ops.layernorm_bf16(x, w, b, 1e-5).float()
    ref = torch.nn.functional.layer_norm(x.double(), (D,), w.double(), b.double(), 1e-5)
    assert (got.double() - ref).abs().max() <= 2 ** -8 * ref.abs().max() + 1e-6      # one bf16 rounding
    got = ops.layernorm_bf16(x, None, None, 1e-5).float()                             # the stack's form: no affine
    ref = torch.nn.functional.layer_norm(x.double(), (D,), None, None, 1e-5)
    assert (got.double() - ref).abs().max() <= 2 ** -8 * ref.abs().max() + 1e-6


@pytest.mark.parametrize("B,G,heads,hd", [(2, 128, 12, 32), (3, 50, 2, 32), (1, 196, 12, 64), (2, 7, 1, 64), (1, 300, 3, 32)])
def test_attention_block(B, G, heads, hd):
    torch.manual_seed(G)
    D = heads * hd
    qkv = (torch.randn(B * G, 3 * D, device=dev()) * 1.5).bfloat16()
    got = ops.attention_bf16(qkv, B, G, heads).float()
    t = qkv.double().reshape(B, G, 3, heads, hd).permute(2, 0, 3, 1, 4)
    att = ((t[0] @ t[1].transpose(-2, -1)) * hd ** -0.5).softmax(-1)
    ref = (att @ t[2]).transpose(1, 2).reshape(B * G, D)
    # probabilities and the output are rounded to bf16 (2^-9 relative each)
    assert (got.double() - ref).abs().max() <= 1.5e-2 * ref.abs().max()
    assert torch.linalg.norm(got.double() - ref) <= 5e-3 * torch.linalg.norm(ref)


@pytest.mark.parametrize("M,K,N", [(256, 384, 1536), (1000, 64, 384), (300, 1600, 384), (130, 384, 64), (128, 768, 2304),
                                   (700, 384, 1600)])
def test_linear_epilogues(M, K, N):
    torch.manual_seed(K + N)
    a = torch.randn(M, K, device=dev()).bfloat16()
    w = (torch.randn(N, K, device=dev()) / K ** 0.5).bfloat16()
    b = torch.randn(N, device=dev())
    acc = a.double() @ w.double().T + b.double()
    if N <= 2048:
        for act, f in ((0, lambda t: t), (1, torch.relu), (2, lambda t: torch.nn.functional.gelu(t))):
            got = ops.linear_bf16_ex(a, w, b, act, 0, None, 0.0, 1.0).double()
            ref = f(acc)
            assert (got - ref).abs().max() <= 2 ** -8 * ref.abs().max() + 1e-3, act
            # elementwise to one bf16 rounding of the fp32 result (GELU: the erfc-exponent polynomial is 2e-5 relative)
            assert ((got - ref).abs() <= 2 ** -8 * ref.abs() + 2e-5 * acc.abs().max()).all(), act
        if N > 64:                                             # [gelu | relu] column split (fc1 + adapter bottleneck GEMM)
            split = (N // 2 + 63) // 64 * 64
            got = ops.linear_bf16_ex(a, w, b, 3, split, None, 0.0, 1.0).double()
            ref = torch.cat([torch.nn.functional.gelu(acc[:, :split]), torch.relu(acc[:, split:])], 1)
            assert ((got - ref).abs() <= 2 ** -8 * ref.abs() + 2e-5 * acc.abs().max()).all()
        res = torch.randn(M, N, device=dev()) * 4
        got = ops.linear_bf16_ex(a, w, b, 0, 0, res, 2.0, 0.7).double()       # general form: the epilogue loads the residual
        ref = 2.0 * res.double() + 0.7 * acc
        assert (got - ref).abs().max() <= 2e-5 * ref.abs().max()
        got = ops.linear_bf16_ex(a, w, b, 0, 0, res, 1.0, 0.7).double()       # in-place form: TMA reduce-add into the stream
        ref = res.double() + 0.7 * acc
        assert (got - ref).abs().max() <= 2e-5 * ref.abs().max()
    else:
        with pytest.raises(RuntimeError):                      # one launch stages at most 2048 bias columns; the stack
            ops.linear_bf16_ex(a, w, b, 0, 0, None, 0.0, 1.0)  # slices wider layers (checked by test_vit_b_width)


def _stack(c, sd):
    blocks = torch.nn.Sequential(*[APFViTLayer(c["D"], c["heads"]) for _ in range(c["depth"])]).eval().to(dev())
    norm = torch.nn.LayerNorm(c["D"]).eval().to(dev())
    head = ClassificationHead(c["D"], c["classes"]).eval().to(dev())
    tsd = synth.to_torch_state(sd)
    blocks.load_state_dict({k[len("blocks."):]: v for k, v in tsd.items() if k.startswith("blocks.")})
    norm.load_state_dict({k[len("encoder_norm."):]: v for k, v in tsd.items() if k.startswith("encoder_norm.")})
    head.load_state_dict({k[len("head."):]: v for k, v in tsd.items() if k.startswith("head.")})
    return blocks, norm, head


@pytest.mark.parametrize("name", list(cases.VIT_CASES))
def test_vit_stack_golden(golden_dir, name):
    c = cases.VIT_CASES[name]
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    sd = synth.apf_vit_state(c["D"], c["depth"], c["classes"], c["seed"])
    tok = synth.vit_tokens(c["B"], c["G"], c["D"], c["seed"])
    blocks, norm, head = _stack(c, sd)
    x, pooled = run_blocks(blocks, to_dev(tok), norm)
    logits = head(pooled)
    ox, op, ol = oracle.apf_vit(sd, tok, c["depth"], c["heads"])
    for got, orc, key in ((x, ox, "x"), (pooled, op, "pooled"), (logits, ol, "logits")):
        assert_tokens_close(got.cpu().numpy(), orc, 1e-2, f"{name}:{key} vs oracle")
        assert_tokens_close(got.cpu().numpy(), g[key], 1e-2, f"{name}:{key} vs reference golden")
    # against a torch model of the kernels' own arithmetic the error is accumulation order + exp2/GELU approximations
    folded = [fold_vit_layer(blk) for blk in blocks]
    ex, ep = vit_bf16_emulation(folded, to_dev(tok), c["heads"], 64, norm.weight.detach(), norm.bias.detach())
    assert rel_err(x.cpu().numpy(), ex.cpu().numpy()) < 4e-3
    assert rel_err(pooled.cpu().numpy(), ep.cpu().numpy()) < 4e-3
    # a single layer called as a module (APFViTLayer.forward) is the first step of the stack
    y1 = blocks[0](to_dev(tok))
    o1 = oracle.apf_vit_layer(sd, "blocks.0.", tok.astype(np.float64), c["heads"])
    assert_tokens_close(y1.cpu().numpy(), o1, 1e-2, f"{name}: single layer")


def test_vit_b_width():
    """ViT-B geometry of the reference's shipped APF config (D = 768: qkv 2304 and fc1 3072 columns are sliced)."""
    c = dict(B=1, G=196, D=768, heads=12, depth=1, classes=15, seed=61)
    sd = synth.apf_vit_state(c["D"], c["depth"], c["classes"], c["seed"])
    tok = synth.vit_tokens(c["B"], c["G"], c["D"], c["seed"])
    blocks, norm, _ = _stack(c, sd)
    x, pooled = run_blocks(blocks, to_dev(tok), norm)
    ox, op, _ = oracle.apf_vit(sd, tok, 1, 12)
    assert_tokens_close(x.cpu().numpy(), ox, 1e-2, "vit-b x")
    assert_tokens_close(pooled.cpu().numpy(), op, 1e-2, "vit-b pooled")


def test_adaptpointformer_end_to_end():
    """AdaptPointFormer.forward (apf.py:348-373): cloud -> logits, against the oracle chained the same way."""
    B, N, G, k, E = 3, 512, 32, 16, 64
    x = synth.make_cloud("clustered", B, N, 71, 3)
    start = synth.start_indices(B, N, 71)
    sd_enc = synth.apf_encoder_state(E, 6, 71)
    sd_vit = synth.apf_vit_state(E, 12, 15, 71)
    tok64, _ = oracle.pointnet_apf(sd_enc, x, start, G, k)
    _, op, ol = oracle.apf_vit(sd_vit, tok64, 12, 2)
    m = AdaptPointFormer(num_classes=15, embedding_dim=E, npoint=G, nsample=k, in_channels=3, precision="fp32")
    for blk in m.blocks:                                     # the reference hard-codes 12 heads (apf.py:312); 64/12 is not
        blk.attention.num_heads = 2                          # a head width, so this small case runs 2 heads of 32
    state = {"point_encoder.encoder." + k_: v for k_, v in synth.to_torch_state(sd_enc).items()}
    state.update(synth.to_torch_state(sd_vit))
    m.load_state_dict(state, strict=True)
    m = m.eval().to(dev())
    feats = m.features(to_dev(x), to_dev(start))
    logits = m(to_dev(x), to_dev(start))
    assert_tokens_close(feats.cpu().numpy(), op, 1e-2, "APF pooled features")
    assert_tokens_close(logits.cpu().numpy(), ol, 1e-2, "APF logits")


def test_vit_errors():
    blk = APFViTLayer(48, 2).eval().to(dev())                # head dim 24: unsupported, must say so
    with pytest.raises(RuntimeError, match="head dim"):
        blk(torch.zeros(1, 4, 48, device=dev()))
    with pytest.raises(RuntimeError):
        APFViTLayer(64, 2).train()(torch.zeros(1, 4, 64))    # train mode (p3tok/train_vit.py) is CUDA-only too
    with pytest.raises(RuntimeError):
        ops.attention_bf16(torch.zeros(8, 192), 2, 4, 2)     # CPU tensor: no fallback


@pytest.mark.parametrize("G", [1, 2, 63, 64, 65, 127, 129])
def test_vit_edge_sequence_lengths(G):
    """One token, block boundaries of the attention kernels (64-key blocks; 128-row persistent form vs the general one)."""
    c = dict(B=2, G=G, D=64, heads=2, depth=1, classes=15, seed=70 + G)
    sd = synth.apf_vit_state(c["D"], 1, 15, c["seed"])
    tok = synth.vit_tokens(c["B"], G, c["D"], c["seed"])
    blocks, norm, _ = _stack(c, sd)
    x, pooled = run_blocks(blocks, to_dev(tok), norm)
    ox, op, _ = oracle.apf_vit(sd, tok, 1, 2)
    assert_tokens_close(x.cpu().numpy(), ox, 1e-2, f"G={G} x")
    assert_tokens_close(pooled.cpu().numpy(), op, 1e-2, f"G={G} pooled")


def test_vit_empty_batch():
    c = dict(B=1, G=4, D=64, heads=2, depth=1, classes=15, seed=69)
    blocks, norm, _ = _stack(c, synth.apf_vit_state(64, 1, 15, 69))
    x, pooled = run_blocks(blocks, torch.zeros(0, 4, 64, device=dev()), norm)
    assert tuple(x.shape) == (0, 4, 64) and tuple(pooled.shape) == (0, 64)


# ------------------------------------------------------------------ Pix4Point's block stack (timm Block variant of next #3)
@pytest.mark.parametrize("name", list(cases.P4P_VIT_CASES))
def test_pointvit_blocks_match_oracle_and_golden(golden_dir, name):
    """p3tok_vit_forward: pre-norm blocks without adapter, the positional embedding re-added in front of every block
    (pix4point.py:254-255), final norm, max over the tokens without the cls row - vs the float64 oracle and the
    torch.nn.TransformerEncoderLayer fixture; bf16 GEMMs / fp32 residual stream: rtol 1e-2."""
    import make_golden
    from p3tok.p4p_model import PointViT, fold_timm_block
    c = cases.P4P_VIT_CASES[name]
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    feats, pos, sd = make_golden.p4p_vit_inputs(c)
    m = PointViT(embed_dim=c["D"], depth=c["depth"], num_heads=c["heads"], k_neighbors=8).eval().to(dev())
    missing = m.load_state_dict(synth.to_torch_state(sd), strict=False)
    assert not missing.unexpected_keys
    params = [t for blk in m.vit.blocks for t in fold_timm_block(blk)]
    out, pooled = ops.vit_blocks(to_dev(feats), to_dev(pos), params, c["heads"], m.norm.weight.detach(), m.norm.bias.detach(), 1e-6, 1)
    ox, og = oracle.pointvit_blocks(sd, feats, pos, c["depth"], c["heads"])
    D = c["D"]
    assert_tokens_close(out.cpu().numpy(), ox, 1e-2, f"{name} normed feats vs oracle")
    assert_tokens_close(out.cpu().numpy(), g["feats"], 1e-2, f"{name} normed feats vs golden")
    assert_tokens_close(pooled.cpu().numpy(), og[:, :D], 1e-2, f"{name} token max vs oracle")
    assert_tokens_close(torch.cat([pooled, out[:, 0]], 1).cpu().numpy(), g["glob"], 1e-2, f"{name} 'max,cls' features vs golden")


def test_pointvit_module_end_to_end():
    """PointViT drop-in (P3Embed -> proj / pos_embed / cls -> blocks -> norm -> 'max,cls'), BASELINE C1 shapes at B = 2, against the
    oracle chain p3embed -> token_head -> pointvit_blocks; state_dict carries the reference's key names."""
    from p3tok.p4p_model import PointViT
    B, N, k, E = 2, 1024, 32, 384
    m = PointViT(embed_dim=E, k_neighbors=k, sample_ratio=1 / 16, precision="fp32").eval().to(dev())
    sd_p = synth.p3embed_state(3, 1 / 16, 4, 4, 256, 7)
    sd_h = synth.token_head_state(256, E, 7)
    sd_v = synth.pointvit_state(E, 12, 7)
    m.patch_embed.load_state_dict(synth.to_torch_state(sd_p), strict=True)
    miss = m.load_state_dict({**synth.to_torch_state(sd_h), **synth.to_torch_state(sd_v)}, strict=False)
    assert not miss.unexpected_keys
    keys = set(m.state_dict())
    assert {"vit.blocks.0.attn.qkv.weight", "vit_blocks.11.mlp.fc2.bias", "vit.norm.weight", "norm.bias", "cls_token", "vit.cls_token",
            "cls_pos", "proj.weight", "pos_embed.2.bias", "patch_embed.convs.1.1.4.running_var"} <= keys
    p = synth.make_cloud("uniform", B, N, 7, 3)
    starts = [synth.start_indices(B, N, 7, 0), synth.start_indices(B, N // 4, 7, 1)]
    glob = m.forward_cls_feat(to_dev(p), None, [to_dev(s) for s in starts])
    p_list, x_list, feats = m(to_dev(p), None, [to_dev(s) for s in starts])
    assert feats.shape == (B, 1 + N // 16, E) and glob.shape == (B, 2 * E)
    op, of, _ = oracle.p3embed(sd_p, p, p.copy(), starts, k, 2)
    hf, hp = oracle.token_head(sd_h, of[-1].astype(np.float32), op[-1])
    ox, og = oracle.pointvit_blocks(sd_v, hf, hp, 12, 6)
    assert_tokens_close(feats.cpu().numpy(), ox, 1e-2, "PointViT feats vs oracle")
    assert_tokens_close(glob.cpu().numpy(), og, 1e-2, "PointViT 'max,cls' vs oracle")

"""Golden-vector cases shared by make_golden.py (build container, imports the reference) and the
tests (anywhere).  Inputs and weights are regenerated from seeds by p3tok.synth, so a fixture
only stores what the REFERENCE returned."""

APF_CASES = {
    # name: dict(kind, B, N, C, G, k, E, seed)
    "apf_uniform": dict(kind="uniform", B=2, N=256, C=3, G=16, k=8, E=64, seed=11),
    "apf_height": dict(kind="clustered", B=2, N=512, C=4, G=24, k=16, E=48, seed=12),
    "apf_dups": dict(kind="duplicates", B=2, N=256, C=3, G=16, k=8, E=32, seed=13),
    "apf_k32": dict(kind="uniform", B=1, N=1024, C=3, G=64, k=32, E=96, seed=14),
    # BASELINE configs[3] ("c4") widths: k = 64 neighbours, E = 384, clustered scene (reduced N / G so the fixture stays small)
    "apf_k64": dict(kind="clustered", B=1, N=8192, C=3, G=128, k=64, E=384, seed=15),
}

P4P_CASES = {
    # name: dict(kind, B, N, k, sample_ratio, embed_dim, seed)   (in_channels=3, scale=4, layers=4)
    "p4p_2stage": dict(kind="uniform", B=2, N=256, k=8, sample_ratio=1.0 / 16, embed_dim=64, seed=21),
    "p4p_1stage": dict(kind="clustered", B=2, N=128, k=16, sample_ratio=0.25, embed_dim=256, seed=22),
    "p4p_2stage_k32": dict(kind="uniform", B=1, N=1024, k=32, sample_ratio=1.0 / 16, embed_dim=256, seed=23),
    # BASELINE configs[2] ("c3") shape for one cloud: 8192 -> 2048 -> 512, k = 32, widths 128 / 256
    "p4p_c3": dict(kind="uniform", B=1, N=8192, k=32, sample_ratio=1.0 / 16, embed_dim=256, seed=24),
}

INDEX_CASES = {
    # bare FPS / kNN at a BASELINE-like shape: name: dict(kind, B, N, G, k, seed)
    "idx_c2like": dict(kind="uniform", B=2, N=2048, G=128, k=32, seed=31),
    "idx_clustered": dict(kind="clustered", B=2, N=1024, G=256, k=32, seed=32),
    "idx_dups": dict(kind="duplicates", B=2, N=512, G=64, k=16, seed=33),
}

FPS_ND_CASES = {
    # farthest_point_sampling on D-dimensional points (pix4point.py:8-53 sums over ALL D coordinates): dict(B, N, G, dims, seed);
    # one index array per D; the dims cover both branches of torch's CPU summation order (below 8 elements / 8 lanes + tail)
    "fps_nd": dict(B=2, N=600, G=80, dims=[1, 2, 4, 5, 6, 7, 8, 9, 11, 12, 15, 16], seed=71),
}

HEAD_CASES = {
    # Pix4Point token head (proj + pos_embed + cls concat): name: dict(B, G, W, E, seed)
    "p4p_head": dict(B=2, G=16, W=64, E=96, seed=41),
}

VIT_CASES = {
    # APF ViT block stack + encoder_norm + max + head on synthetic tokens: name: dict(B, G, D, heads, depth, classes, seed)
    "vit_small": dict(B=2, G=50, D=64, heads=2, depth=2, classes=15, seed=51),      # head dim 32, ragged G
    "vit_hd64": dict(B=1, G=70, D=128, heads=2, depth=3, classes=15, seed=52),      # head dim 64
    "vit_s12": dict(B=1, G=128, D=384, heads=12, depth=12, classes=15, seed=53),    # ViT-S geometry of BASELINE config 2
}

P4P_VIT_CASES = {
    # Pix4Point's ViT tail (pix4point.py:254-271): timm-style pre-norm blocks with the positional embedding re-added before every
    # block, final norm, 'max,cls' features: dict(B, G, D, heads, depth, seed)  (sequence = 1 cls row + G tokens)
    "p4p_vit": dict(B=2, G=20, D=128, heads=2, depth=3, seed=61),          # head dim 64, ragged sequence of 21
    "p4p_vit_s": dict(B=1, G=64, D=384, heads=6, depth=12, seed=62),       # vit_small_patch16_384 geometry, BASELINE C1 token count
}

P4P_TRAIN_CASES = {
    # training-mode P3Embed (1 stage) through the reference's own forward + autograd: dict(B, N, k, W, seed)
    "p4p_train": dict(B=2, N=64, k=8, W=32, seed=96),
}

VIT_TRAIN_CASES = {
    # backward through the frozen APFViTLayer stack + trainable encoder_norm + token max: dict(B, G, D, heads, depth, seed)
    "vit_train": dict(B=2, G=20, D=64, heads=2, depth=2, seed=97),
}

P4P_VIT_TRAIN_CASES = {
    # Pix4Point's block loop + final norm + 'max,cls' features under autograd, every parameter, feats and pos asking for a gradient
    "p4p_vit_train": dict(B=3, G=9, D=64, heads=2, depth=2, seed=151),
}

VIT_FULL_TRAIN_CASES = {
    # the whole token consumer in TRAIN mode under autograd - blocks (every parameter asks for a gradient), encoder_norm, max
    # over tokens, dropout, ClassificationHead with batch-statistics BatchNorm1d: dict(B, G, D, heads, depth, classes, seed,
    # p_adapter, p_pool, p_head) - the dropout rates become forced keep masks (0 = dropout off)
    "vit_train_full": dict(B=4, G=12, D=64, heads=2, depth=2, classes=15, seed=131, p_adapter=0.0, p_pool=0.0, p_head=0.0),
    "vit_train_masked": dict(B=6, G=9, D=64, heads=4, depth=2, classes=7, seed=137, p_adapter=0.25, p_pool=0.1, p_head=0.4),
}

TRAIN_CASES = {
    # training-mode APF Encoder (batch-statistics BN) + autograd through both max-pools: dict(B, N, C, G, k, E, seed)
    "apf_train": dict(B=2, N=128, C=3, G=6, k=8, E=32, seed=95),
}

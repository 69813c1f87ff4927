"""Golden parity cases shared by the generator (make_golden.py, needs /root/reference) and the tests.

Each op-level case names a reference module class, its constructor arguments and the input geometry.
The generator instantiates the *reference* class, fills its BatchNorms / heads with
hyperseg_b200.synthetic.deterministic_init, runs it on seeded inputs and stores inputs + outputs in
tests/golden/ops.npz.  Tests rebuild the same module from this package (same state_dict keys -> same
seeded values) and compare.
"""
import torch

# kind:
#   nopad      HyperPatchNoPadding(Cin, Cout, 1, groups=g)                     hyperseg_v1_0.py:455
#   block1x1   make_hyper_patch_conv2d_block(Cin, Cout, 1) = conv + BN + ReLU  hyperseg_v1_0.py:728
#   ir         HyperPatchInvertedResidual(Cin, Cout, 3, expand_ratio=e)        hyperseg_v1_0.py:281
#   hpconv     HyperPatchConv2d(Cin, Cout, k, padding=p, ...)                  hyperseg_v1_0.py:560
#   metapatch  MetaPatchConv2d(Cin, Cout, k, padding=p, dilation=d, groups=g, padding_mode=m)   meta_patch.py:60
#   mpblock    make_meta_patch_conv2d_block(...) = MetaPatchConv2d + BN + act  meta_patch.py:228
#   metaconv   MetaConv2d(Cin, Cout, k, padding=p, dilation=d, groups=g, padding_mode=m)        meta_conv.py:9
#   v01_ir     hyperseg_v0_1.HyperPatchInvertedResidual(Cin, Cout, 3, expand_ratio=e)           hyperseg_v0_1.py:205
# `head`: (signal_total_channels, signal_channels, signal_index, groups) -> init_signal2weights is called and the
#         second forward argument is the signal map instead of ready-made weights.
OP_CASES = {
    # -- 1x1 patch conv --------------------------------------------------------------------------------
    "nopad_basic": dict(kind="nopad", B=2, Cin=6, Cout=8, H=8, W=12, fh=2, fw=3),
    "nopad_p1": dict(kind="nopad", B=2, Cin=82, Cout=64, H=2, W=4, fh=2, fw=4),          # HyperSeg-M level 0 shape
    "nopad_p2": dict(kind="nopad", B=1, Cin=94, Cout=32, H=4, W=8, fh=2, fw=4),          # level 1 shape
    "nopad_p4": dict(kind="nopad", B=1, Cin=44, Cout=16, H=8, W=16, fh=2, fw=4),         # level 2 shape
    "nopad_groups": dict(kind="nopad", B=1, Cin=8, Cout=12, H=8, W=8, fh=2, fw=2, groups=4),
    "nopad_odd": dict(kind="nopad", B=3, Cin=5, Cout=3, H=6, W=9, fh=2, fw=3),           # odd Cin/Cout, 3x3 patches
    "nopad_ragged": dict(kind="nopad", B=1, Cin=10, Cout=6, H=3, W=35, fh=1, fw=7),      # 7 patches of 3x5
    "nopad_single": dict(kind="nopad", B=1, Cin=4, Cout=4, H=5, W=7, fh=1, fw=1),        # one patch = whole map
    "block1x1": dict(kind="block1x1", B=2, Cin=12, Cout=10, H=8, W=8, fh=2, fw=2),
    "block1x1_head": dict(kind="block1x1", B=2, Cin=12, Cout=10, H=8, W=8, fh=2, fw=2, head=(48, 32, 8, 8)),
    "nopad_head_pad": dict(kind="nopad", B=1, Cin=7, Cout=5, H=4, W=6, fh=2, fw=3, head=(24, 24, 0, 4)),  # hp=35 -> 36
    # -- fused inverted residual ---------------------------------------------------------------------------
    "ir_small": dict(kind="ir", B=2, Cin=6, Cout=5, expand=2, H=12, W=16, fh=3, fw=4),
    "ir_res": dict(kind="ir", B=1, Cin=8, Cout=8, expand=2, H=8, W=8, fh=2, fw=2),       # use_res_connect
    "ir_L3": dict(kind="ir", B=1, Cin=24, Cout=16, expand=2, H=16, W=24, fh=2, fw=3),    # HyperSeg-M level 3, 8x8 patches
    "ir_L4": dict(kind="ir", B=1, Cin=34, Cout=19, expand=2, H=32, W=32, fh=2, fw=2),    # level 4, 16x16 patches
    "ir_1patch": dict(kind="ir", B=2, Cin=4, Cout=3, expand=3, H=6, W=10, fh=1, fw=1),   # all four borders reflect
    "ir_rect": dict(kind="ir", B=1, Cin=5, Cout=7, expand=1, H=6, W=20, fh=3, fw=2),     # 2x10 patches, expand 1
    "ir_head": dict(kind="ir", B=2, Cin=6, Cout=5, expand=2, H=8, W=8, fh=2, fw=2, head=(64, 32, 16, 4)),
    "ir_L4_head": dict(kind="ir", B=1, Cin=34, Cout=19, expand=2, H=32, W=32, fh=2, fw=2, head=(1280, 320, 0, 4)),
    # -- generic patch conv / MetaPatch ----------------------------------------------------------------------
    "hpconv_k3": dict(kind="hpconv", B=2, Cin=4, Cout=6, k=3, pad=1, H=8, W=12, fh=2, fw=3),
    "hpconv_k3_head": dict(kind="hpconv", B=1, Cin=4, Cout=6, k=3, pad=1, H=8, W=8, fh=2, fw=2, head=(32, 16, 8, 2)),
    "metapatch_pw": dict(kind="metapatch", B=2, Cin=6, Cout=4, k=1, pad=0, H=8, W=8, fh=2, fw=2),
    "metapatch_dw": dict(kind="metapatch", B=2, Cin=6, Cout=6, k=3, pad=1, groups=6, H=8, W=12, fh=2, fw=3),
    "metapatch_dil": dict(kind="metapatch", B=1, Cin=3, Cout=5, k=3, pad=2, dil=2, mode="replicate", H=12, W=12, fh=2, fw=3),
    "metapatch_circ": dict(kind="metapatch", B=1, Cin=4, Cout=4, k=3, pad=1, groups=2, mode="circular", H=6, W=9, fh=2, fw=3),
    "metapatch_1patch": dict(kind="metapatch", B=1, Cin=3, Cout=3, k=3, pad=1, H=5, W=6, fh=1, fw=1),
    "mpblock_dw": dict(kind="mpblock", B=2, Cin=8, Cout=8, k=3, groups=8, act="relu6", H=8, W=8, fh=2, fw=2),
    "mpblock_pw_lin": dict(kind="mpblock", B=2, Cin=8, Cout=5, k=1, act=None, H=8, W=8, fh=2, fw=2),
    "v01_ir": dict(kind="v01_ir", B=2, Cin=6, Cout=5, expand=2, H=8, W=12, fh=2, fw=3),
    "v01_ir_res": dict(kind="v01_ir", B=1, Cin=6, Cout=6, expand=2, H=8, W=8, fh=2, fw=2),
    "v01_ir_e1": dict(kind="v01_ir", B=1, Cin=6, Cout=4, expand=1, H=8, W=8, fh=2, fw=2),
    # -- per-sample dynamic conv ---------------------------------------------------------------------------------
    "metaconv_zeros": dict(kind="metaconv", B=4, Cin=3, Cout=3, k=3, pad=1, groups=3, mode="zeros", H=9, W=7),
    "metaconv_reflect": dict(kind="metaconv", B=2, Cin=4, Cout=6, k=3, pad=1, groups=2, mode="reflect", H=6, W=8),
    "metaconv_valid": dict(kind="metaconv", B=2, Cin=5, Cout=2, k=(1, 3), pad=0, mode="zeros", H=4, W=9),
}

# head-only cases: (signal_total, signal_channels, signal_index, groups, hyper_params) -> apply_signal2weights
HEAD_CASES = {
    "head_M_L0": dict(B=2, C=1280, sc=416, idx=0, groups=32, hp=5248, fh=2, fw=3),
    "head_M_L4": dict(B=1, C=1280, sc=320, idx=0, groups=4, hp=4216, fh=2, fw=2),
    "head_pad": dict(B=2, C=96, sc=48, idx=24, groups=8, hp=203, fh=3, fw=2),            # out_ch 208, 203 used
    "head_unify": dict(B=1, C=1280, sc=512, idx=768, groups=16, hp=3676, fh=1, fw=2),    # S-Cityscapes shared head
}

# whole-model cases: config name (hyperseg_b200.synthetic.CONFIGS), batch, height, width
MODEL_CASES = {
    "model_m_128x256": dict(config="hyperseg-m", B=1, H=128, W=256),         # BASELINE.json configs[0]
    "model_m_b2_64x128": dict(config="hyperseg-m", B=2, H=64, W=128),
    "model_s_camvid_128x192": dict(config="hyperseg-s-camvid", B=1, H=128, W=192),
    "model_l_camvid_64x128": dict(config="hyperseg-l-camvid", B=1, H=64, W=128),
    "model_s_city_128x192": dict(config="hyperseg-s-cityscapes", B=1, H=128, W=192),
    "model_l_voc_128x128": dict(config="hyperseg-l-voc", B=2, H=128, W=128),
}
MODEL_STRIDE = 2     # logits are stored at this spatial stride

DIVIDE_CASES = [
    (1280, [5248, 3008, 704, 2352, 4216], 32),
    (1280, [4160, 992, 208, 3676], 32),
    (1280, [5248, 3008, 704, 2352, 1892], 64),
    (1280, [5248, 3008, 704, 2352, 2352, 1892], 64),
    (1280, [100, 100, 100, 50], 8),
    (512, [7, 7, 9, 300, 300, 1], 4),
    (64, [10], 8),
    (256, [3, 1, 2], 16),
]


# gradient cases (training path): op cases whose inputs get requires_grad; loss = sum(y * r) with a seeded r
GRAD_CASES = ["nopad_basic", "nopad_groups", "nopad_odd", "nopad_p1", "metapatch_pw", "metapatch_dw", "metapatch_dil",
              "metapatch_circ", "metapatch_1patch", "hpconv_k3", "hpconv_k3_head", "block1x1_head", "mpblock_dw", "v01_ir_res",
              "ir_small", "ir_res", "ir_head", "ir_rect"]
# the same, with the module in train() mode (batch-statistics BatchNorm inside the block): y, dx, dw and the
# updated running statistics of every BatchNorm are stored
TRAIN_OP_CASES = ["ir_small", "ir_res", "ir_1patch", "ir_head", "block1x1", "mpblock_dw", "v01_ir"]
# whole-model training steps at sizes the CPU finishes quickly: BASELINE config 4's model (HyperSeg-L VOC, hyperseg_v0_1)
# and the north-star model (HyperSeg-M, hyperseg_v1_0: heads inside the layers, stage-wise inverted-residual blocks)
TRAIN_CASE = dict(config="hyperseg-l-voc", B=2, H=128, W=128, seed=5)
TRAIN_PARAMS = ["weight_mapper.out_conv.conv_0.weight", "weight_mapper.out_conv.conv_5.weight", "weight_mapper.flat_0.0.weight",
                "weight_mapper.down_0.0.weight", "decoder.level_0.0.1.weight", "decoder.level_2.0.conv.1.1.bias",
                "decoder.level_5.0.conv.2.1.weight", "backbone._conv_stem.weight", "backbone._blocks.5._project_conv.weight",
                "backbone._conv_head.weight"]
TRAIN_CASE_V10 = dict(config="hyperseg-m", B=2, H=128, W=256, seed=7)
TRAIN_PARAMS_V10 = ["decoder.level_0.0.0.signal2weights.weight", "decoder.level_2.0.0.signal2weights.weight",
                    "decoder.level_3.0.signal2weights.weight", "decoder.level_4.0.signal2weights.weight",
                    "decoder.level_1.0.1.weight", "decoder.level_3.0.bn2.weight", "decoder.level_4.0.bn3.bias",
                    "weight_mapper.in_conv.0.weight", "weight_mapper.in_conv.1.weight",
                    "backbone._conv_stem.weight", "backbone._feat_fc_4.0.weight", "backbone._conv_head.weight"]
TRAIN_STEPS = {"train": (TRAIN_CASE, TRAIN_PARAMS, "decoder.level_0.0.1.running_mean"),
               "train_v10": (TRAIN_CASE_V10, TRAIN_PARAMS_V10, "decoder.level_4.0.bn1.running_mean")}


def grad_probe(name, shape):
    g = torch.Generator().manual_seed(case_seed(name) + 17)
    return torch.randn(*shape, generator=g)


def train_labels(case, num_classes):
    g = torch.Generator().manual_seed(case["seed"])
    lab = torch.randint(0, num_classes, (case["B"], case["H"], case["W"]), generator=g)
    lab[:, :4, :] = 255            # some ignored pixels
    return lab


def case_seed(name):
    import zlib
    return zlib.crc32(name.encode()) % (2 ** 31 - 1)


def op_inputs(name, case, hyper_params):
    """Seeded (x, second forward argument) for an op case; the second argument is either ready-made
    per-patch weights ~N(0, 0.3^2) (the distribution of SURVEY.md section 8d) or, with `head`, a signal map."""
    g = torch.Generator().manual_seed(case_seed(name))
    x = torch.randn(case["B"], case["Cin"], case["H"], case["W"], generator=g)
    if case["kind"] == "metaconv":
        w = torch.randn(case["B"], int(hyper_params), generator=g) * 0.3
    elif "head" in case:
        w = torch.randn(case["B"], case["head"][0], case["fh"], case["fw"], generator=g).abs() * 0.5
    else:
        w = torch.randn(case["B"], int(hyper_params), case["fh"], case["fw"], generator=g) * 0.3
    return x, w


def head_inputs(name, case):
    g = torch.Generator().manual_seed(case_seed(name))
    return torch.randn(case["B"], case["C"], case["fh"], case["fw"], generator=g)


def build_op_module(ns, case):
    """Instantiate the module of an op case from namespace `ns`, which provides the classes (either the
    reference's modules or hyperseg_b200.nn's mirror)."""
    import torch.nn as nn
    kind = case["kind"]
    k = case.get("k", 1)
    pad = case.get("pad", 0)
    if kind == "nopad":
        m = ns["HyperPatchNoPadding"](case["Cin"], case["Cout"], 1, groups=case.get("groups", 1))
    elif kind == "block1x1":
        m = ns["make_hyper_patch_conv2d_block"](case["Cin"], case["Cout"], 1)
    elif kind == "ir":
        m = ns["HyperPatchInvertedResidual"](case["Cin"], case["Cout"], 3, expand_ratio=case["expand"])
    elif kind == "hpconv":
        m = ns["HyperPatchConv2d"](case["Cin"], case["Cout"], k, padding=pad)
    elif kind == "metapatch":
        m = ns["MetaPatchConv2d"](case["Cin"], case["Cout"], k, padding=pad, dilation=case.get("dil", 1),
                                  groups=case.get("groups", 1), padding_mode=case.get("mode", "reflect"))
    elif kind == "mpblock":
        act = {"relu6": nn.ReLU6(inplace=True), "relu": nn.ReLU(True), None: None}[case.get("act", "relu")]
        m = ns["make_meta_patch_conv2d_block"](case["Cin"], case["Cout"], k, groups=case.get("groups", 1),
                                               act_layer=act)
    elif kind == "metaconv":
        m = ns["MetaConv2d"](case["Cin"], case["Cout"], k, padding=pad, dilation=case.get("dil", 1),
                             groups=case.get("groups", 1), padding_mode=case.get("mode", "zeros"))
    elif kind == "v01_ir":
        m = ns["V01HyperPatchInvertedResidual"](case["Cin"], case["Cout"], 3, expand_ratio=case["expand"])
    else:
        raise KeyError(kind)
    if "head" in case:
        _, sc, idx, groups = case["head"]
        owner = m[0] if isinstance(m, nn.Sequential) else m
        owner.init_signal2weights(sc, idx, groups)
    return m

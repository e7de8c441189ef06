"""Parameter table and named workloads of the UniMatch matching path.

The drop-in module keeps the reference's `state_dict` layout (reference
`unimatch/unimatch.py:34-62`, `backbone.py:49-86`, `transformer.py:22-40,214-220`,
`attention.py:177-178`, `reg_refine.py:10-12,29-35,62-66,92-104`) so that reference checkpoints
load unchanged.  Here the layout is a flat table key -> shape; the module tree is generated from it.
"""
from collections import OrderedDict


def param_spec(num_scales=1, feature_channels=128, upsample_factor=8, num_head=1, ffn_dim_expansion=4,
               num_transformer_layers=6, reg_refine=False, task="flow"):
    """Ordered {state_dict key: shape}, in the reference's registration order."""
    C = feature_channels
    p = OrderedDict()

    def conv(name, cout, cin, kh, kw, bias):
        p[name + ".weight"] = (cout, cin, kh, kw)
        if bias:
            p[name + ".bias"] = (cout,)

    # backbone.py:49-86 (InstanceNorm has no parameters; downsample.0 is a biased 1x1 conv)
    conv("backbone.conv1", 64, 3, 7, 7, False)
    cin = 64
    for li, dim in ((1, 64), (2, 96), (3, 128)):
        for bi in range(2):
            pf = "backbone.layer%d.%d" % (li, bi)
            conv(pf + ".conv1", dim, cin, 3, 3, False)
            conv(pf + ".conv2", dim, dim, 3, 3, False)
            if bi == 0 and cin != dim:
                conv(pf + ".downsample.0", dim, cin, 1, 1, True)
            cin = dim
    conv("backbone.conv2", C, 128, 1, 1, True)
    if num_scales > 1:
        conv("backbone.trident_conv", C, C, 3, 3, False)

    # transformer.py:22-40
    for i in range(num_transformer_layers):
        for layer, ffn in (("self_attn", False), ("cross_attn_ffn", True)):
            pf = "transformer.layers.%d.%s" % (i, layer)
            for nm in ("q_proj", "k_proj", "v_proj", "merge"):
                p[pf + "." + nm + ".weight"] = (C, C)
            p[pf + ".norm1.weight"] = (C,)
            p[pf + ".norm1.bias"] = (C,)
            if ffn:
                p[pf + ".mlp.0.weight"] = (2 * C * ffn_dim_expansion, 2 * C)
                p[pf + ".mlp.2.weight"] = (C, 2 * C * ffn_dim_expansion)
                p[pf + ".norm2.weight"] = (C,)
                p[pf + ".norm2.bias"] = (C,)

    # attention.py:177-178
    for nm in ("q_proj", "k_proj"):
        p["feature_flow_attn.%s.weight" % nm] = (C, C)
        p["feature_flow_attn.%s.bias" % nm] = (C,)

    # unimatch.py:47-52
    if (not reg_refine) or task == "depth":
        conv("upsampler.0", 256, 2 + C, 3, 3, True)
        conv("upsampler.2", upsample_factor ** 2 * 9, 256, 1, 1, True)

    # unimatch.py:54-62, reg_refine.py
    if reg_refine:
        fd = 2 if task == "flow" else 1
        conv("refine_proj", 256, 128, 1, 1, True)
        conv("refine.encoder.convc1", 256, 81, 1, 1, True)
        conv("refine.encoder.convc2", 192, 256, 3, 3, True)
        conv("refine.encoder.convf1", 128, fd, 7, 7, True)
        conv("refine.encoder.convf2", 64, 128, 3, 3, True)
        conv("refine.encoder.conv", 128 - fd, 256, 3, 3, True)
        for g in ("z", "r", "q"):
            conv("refine.gru.conv%s1" % g, 128, 384, 1, 5, True)
        for g in ("z", "r", "q"):
            conv("refine.gru.conv%s2" % g, 128, 384, 5, 1, True)
        # registration order in the reference is z1 r1 q1 z2 r2 q2 (reg_refine.py:29-35)
        conv("refine.flow_head.conv1", 256, 128, 3, 3, True)
        conv("refine.flow_head.conv2", fd, 256, 3, 3, True)
        if task != "depth":
            conv("refine.mask.0", 256, 128, 3, 3, True)
            conv("refine.mask.2", upsample_factor ** 2 * 9, 256, 1, 1, True)
    return p


# BASELINE.json `configs`, with the flags the reference's scripts pass
# (scripts/gmflow_evaluate.sh:26-39, scripts/gmstereo_evaluate.sh:11-22, scripts/gmdepth_evaluate.sh:12-19).
WORKLOADS = {
    "gmflow-scale1": dict(
        model=dict(num_scales=1, upsample_factor=8, reg_refine=False, task="flow"),
        call=dict(attn_type="swin", attn_splits_list=[2], corr_radius_list=[-1], prop_radius_list=[-1],
                  num_reg_refine=1, task="flow"),
        pad=16),
    "gmflow-scale2": dict(
        model=dict(num_scales=2, upsample_factor=4, reg_refine=False, task="flow"),
        call=dict(attn_type="swin", attn_splits_list=[2, 8], corr_radius_list=[-1, 4], prop_radius_list=[-1, 1],
                  num_reg_refine=1, task="flow"),
        pad=32),
    "gmflow-scale2-regrefine6": dict(
        model=dict(num_scales=2, upsample_factor=4, reg_refine=True, task="flow"),
        call=dict(attn_type="swin", attn_splits_list=[2, 8], corr_radius_list=[-1, 4], prop_radius_list=[-1, 1],
                  num_reg_refine=6, task="flow"),
        pad=32),
    "gmstereo-scale2": dict(
        model=dict(num_scales=2, upsample_factor=4, reg_refine=False, task="stereo"),
        call=dict(attn_type="self_swin2d_cross_swin1d", attn_splits_list=[2, 8], corr_radius_list=[-1, 4],
                  prop_radius_list=[-1, 1], num_reg_refine=1, task="stereo"),
        pad=32),
    "gmstereo-scale2-regrefine3": dict(
        model=dict(num_scales=2, upsample_factor=4, reg_refine=True, task="stereo"),
        call=dict(attn_type="self_swin2d_cross_swin1d", attn_splits_list=[2, 8], corr_radius_list=[-1, 4],
                  prop_radius_list=[-1, 1], num_reg_refine=3, task="stereo"),
        pad=32),
    "gmdepth-scale1": dict(
        model=dict(num_scales=1, upsample_factor=8, reg_refine=False, task="depth"),
        call=dict(attn_type="swin", attn_splits_list=[2], prop_radius_list=[-1], num_reg_refine=1, task="depth",
                  min_depth=1.0 / 10, max_depth=1.0 / 0.5, num_depth_candidates=64),
        pad=16),
    "gmdepth-scale1-regrefine1": dict(
        model=dict(num_scales=1, upsample_factor=8, reg_refine=True, task="depth"),
        call=dict(attn_type="swin", attn_splits_list=[2], prop_radius_list=[-1], num_reg_refine=1, task="depth",
                  min_depth=1.0 / 10, max_depth=1.0 / 0.5, num_depth_candidates=64),
        pad=16),
}

# BASELINE.json configs[i] -> (workload, per-job batch, H, W)
BASELINE_CONFIGS = [
    ("gmflow-scale1", 1, 256, 320),
    ("gmflow-scale1", 32, 480, 832),
    ("gmstereo-scale2", 16, 544, 960),
    ("gmflow-scale2-regrefine6", 64, 480, 832),
    ("gmdepth-scale1-regrefine1", 64, 384, 512),
]

"""Generate tests/golden/golden.pt by running the REFERENCE (build container only).

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py

Imports `/root/reference/unimatch` read-only, loads the synthetic weights into the reference modules,
runs every case of `cases.py` through the reference's own functions / `UniMatch.forward`, asserts the
oracle restatement agrees (bit-exact where the same ATen ops run in the same order), and stores the
reference outputs (fp32) for the travelling test `tests/test_oracle_golden.py`.
The reference has no tests or fixtures of its own, so these vectors are the pin.
"""
import os
import sys
import warnings

import torch

warnings.filterwarnings("ignore")
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, "/root/reference")

import cases  # noqa: E402
from cases import O  # noqa: E402

from unimatch import attention as R_att, geometry as R_geo, matching as R_mat, utils as R_utl  # noqa: E402
from unimatch.unimatch import UniMatch as RefUniMatch  # noqa: E402

from unimatch_b200.synthetic import synthetic_state_dict  # noqa: E402


def ref_model(model_kwargs, sd):
    m = RefUniMatch(**model_kwargs).eval()
    missing = m.load_state_dict(sd, strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    return m


_models = {}


def model_for(model_kwargs, sd_key="op"):
    key = (sd_key, tuple(sorted(model_kwargs.items())))
    if key not in _models:
        _models[key] = ref_model(model_kwargs, synthetic_state_dict(seed=326, **model_kwargs))
    return _models[key]


def mask2d(x, k):
    h, w = x["h"], x["w"]
    return R_utl.generate_shift_window_attn_mask((h, w), h // k, w // k, h // k // 2, w // k // 2, device="cpu")


REF_OPS = {
    "attn_full": lambda x, m: R_att.single_head_full_attention(x["q"], x["k"], x["v"]),
    "attn_full_1d": lambda x, m: R_att.single_head_full_attention_1d(x["q"], x["k"], x["v"], h=x["h"], w=x["w"]),
    "attn_window_2d": lambda x, m: R_att.single_head_split_window_attention(
        x["q"], x["k"], x["v"], num_splits=2, with_shift=False, h=x["h"], w=x["w"]),
    "attn_window_2d_shift": lambda x, m: R_att.single_head_split_window_attention(
        x["q"], x["k"], x["v"], num_splits=2, with_shift=True, h=x["h"], w=x["w"], attn_mask=mask2d(x, 2)),
    "attn_window_2d_shift_k4": lambda x, m: R_att.single_head_split_window_attention(
        x["q"], x["k"], x["v"], num_splits=4, with_shift=True, h=x["h"], w=x["w"], attn_mask=mask2d(x, 4)),
    "attn_window_1d_shift": lambda x, m: R_att.single_head_split_window_attention_1d(
        x["q"], x["k"], x["v"], num_splits=4, with_shift=True, h=x["h"], w=x["w"],
        attn_mask=R_utl.generate_shift_window_attn_mask_1d(24, 6, 3, device="cpu")),
    "attn_window_1d": lambda x, m: R_att.single_head_split_window_attention_1d(
        x["q"], x["k"], x["v"], num_splits=4, with_shift=False, h=x["h"], w=x["w"]),
    "add_position_k2": lambda x, m: torch.stack(R_utl.feature_add_position(x["f0"], x["f1"], 2, 128)),
    "add_position_k1": lambda x, m: torch.stack(R_utl.feature_add_position(x["f0"], x["f1"], 1, 128)),
    "transformer_swin_k2": lambda x, m: torch.stack(m.transformer(x["f0"], x["f1"], attn_type="swin", attn_num_splits=2)),
    "transformer_swin_k1": lambda x, m: torch.stack(m.transformer(x["f0"], x["f1"], attn_type="swin", attn_num_splits=1)),
    "transformer_stereo_k2": lambda x, m: torch.stack(
        m.transformer(x["f0"], x["f1"], attn_type="self_swin2d_cross_swin1d", attn_num_splits=2)),
    "transformer_stereo_k1": lambda x, m: torch.stack(
        m.transformer(x["f0"], x["f1"], attn_type="self_swin2d_cross_1d", attn_num_splits=1)),
    "global_corr": lambda x, m: R_mat.global_correlation_softmax(x["f0"], x["f1"], False)[0],
    "global_corr_bidir": lambda x, m: R_mat.global_correlation_softmax(x["f0"], x["f1"], True)[0],
    "local_corr_r4": lambda x, m: R_mat.local_correlation_softmax(x["f0"], x["f1"], 4)[0],
    "local_corr_volume": lambda x, m: R_mat.local_correlation_with_flow(x["f0"], x["f1"], x["flow"], 4),
    "global_corr_stereo": lambda x, m: R_mat.global_correlation_softmax_stereo(x["f0"], x["f1"])[0],
    "local_corr_stereo_r4": lambda x, m: R_mat.local_correlation_softmax_stereo(x["f0"], x["f1"], 4)[0],
    "depth_corr": lambda x, m: R_mat.correlation_softmax_depth(x["f0"], x["f1"], x["K"], x["pose"], x["cand"])[0],
    "depth_corr_bidir_argmax": lambda x, m: R_mat.correlation_softmax_depth(
        x["f0"], x["f1"], x["K"], x["pose"], x["cand"], depth_from_argmax=True, pred_bidir_depth=True)[0],
    "flow_warp": lambda x, m: R_geo.flow_warp(x["f1"], x["flow"]),
    "rigid_flow": lambda x, m: R_geo.compute_flow_with_depth_pose(1.0 / x["cand"][:, 3], x["K"], extrinsics_rel=x["pose"]),
    "prop_global": lambda x, m: m.feature_flow_attn(x["f0"], x["flow"], local_window_attn=False),
    "prop_local_r1": lambda x, m: m.feature_flow_attn(x["f0"], x["flow"], local_window_attn=True, local_window_radius=1),
    "prop_local_r1_disp": lambda x, m: m.feature_flow_attn(x["f0"], x["flow"], local_window_attn=True,
                                                           local_window_radius=1),
    "convex_upsample": lambda x, m: R_utl.upsample_flow_with_mask(x["flow"], x["mask"], 4),
    "update_block_flow": lambda x, m: torch.cat(list(m.refine(x["net"], x["inp"], x["corr"], x["flow"])), dim=1),
    "backbone_s2": lambda x, m: torch.cat([t.flatten(1) for t in m.backbone(x["img"])], dim=1),
}


def main():
    torch.manual_seed(0)
    golden = {}
    report = []
    with torch.no_grad():
        for name, (make_in, run_oracle) in cases.OP_CASES.items():
            mk = cases.OP_CASE_WEIGHTS[name]
            m = model_for(mk)
            sd = synthetic_state_dict(seed=326, **mk)
            x = make_in()
            ref = REF_OPS[name](x, m).float().contiguous()
            got = run_oracle(make_in(), sd)
            err = (ref - got).abs().max().item()
            scale = ref.abs().max().item()
            report.append((name, tuple(ref.shape), err, scale))
            assert err <= 1e-5 * max(scale, 1.0), (name, err, scale)
            golden[name] = ref.clone()
        for name in cases.E2E_CASES:
            cfg, sd, batch, call = cases.e2e_setup(name)
            m = ref_model(cfg["model"], sd)
            out = m(batch["img0"], batch["img1"], intrinsics=batch.get("intrinsics"), pose=batch.get("pose"), **call)
            assert len(out["flow_preds"]) == 1
            ref = out["flow_preds"][-1].float().contiguous()
            got = cases.e2e_oracle(name)
            err = (ref - got).abs().max().item()
            scale = ref.abs().max().item()
            report.append((name, tuple(ref.shape), err, scale))
            assert err <= 1e-4 * max(scale, 1.0), (name, err, scale)
            golden[name] = ref.clone()
    for r in report:
        print("%-28s shape=%-22s max|ref-oracle|=%.3e  max|ref|=%.3e" % (r[0], r[1], r[2], r[3]))
    path = os.path.join(HERE, "golden.pt")
    torch.save({"torch": str(torch.__version__), "threads": torch.get_num_threads(), "vectors": golden}, path)
    print("wrote", path, "%.2f MB" % (os.path.getsize(path) / 1e6))


if __name__ == "__main__":
    main()

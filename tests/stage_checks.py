"""Teacher-forced stage parity of `unimatch_b200.UniMatch` against the oracle (== reference, tests/golden).

The oracle's forward records its intermediate tensors (`taps`: encoder features, warped / transformed features, flow after
correlation and after propagation, flow after every refinement iteration, the upsampling mask).  Each stage of the module
(`UniMatch._stage_*`) is then run ON THE ORACLE'S INPUTS for that stage and compared with the oracle's output for it, so an
error cannot hide behind -- or be blamed on -- the amplification of earlier stages (unimatch/unimatch.py:136-354).

Used at the BASELINE shape on the GPU (tests/test_stages_gpu.py, the CUDA kernels) and at a small shape on the CPU
(tests/test_stages_cpu.py, oracle-backed kernels of tests/refops.py: checks this harness and the host orchestration).

Tolerances (stated per stage, asserted):
  * feature stages  : max |diff| <= FEAT_TOL x max |ref|
  * flow stages     : mean EPE <= FLOW_MEAN px and max EPE <= FLOW_MAX px (at the stage's own resolution)
"""
import torch

from oracle import unimatch_oracle as O
from unimatch_b200 import UniMatch
from unimatch_b200.spec import WORKLOADS
from unimatch_b200.synthetic import BENCH_WEIGHTS, synthetic_batch, synthetic_state_dict

FEAT_TOL = 3e-5          # encoder, warp: one pass of fp32-faithful arithmetic
TRANSFORMER_TOL = 1e-4   # six blocks (24 GEMMs, 12 attention calls, 18 LayerNorms) in sequence
FLOW_MEAN, FLOW_MAX = 2e-4, 5e-3     # px, per stage, teacher-forced
E2E_MEAN, E2E_MAX = 1e-2, 1e-1       # px at full resolution, free-running end to end (= bench.py's tolerance)


def cl(t, dev):
    """oracle NCHW -> channel-last on `dev`"""
    return t.permute(0, 2, 3, 1).contiguous().to(dev)


def feat_err(got, ref_nchw):
    ref = ref_nchw.permute(0, 2, 3, 1)
    got = got.detach().float().cpu().reshape(ref.shape)
    return (got - ref).abs().max().item() / max(ref.abs().max().item(), 1e-30)


def flow_err(got_cl, ref_nchw):
    ref = ref_nchw.permute(0, 2, 3, 1)
    d = (got_cl.detach().float().cpu().reshape(ref.shape) - ref).norm(dim=-1)
    return d.mean().item(), d.max().item()


def run(dev, workload="gmflow-scale2-regrefine6", H=480, W=832, weights=None, report=print):
    """Returns {stage: error}; raises AssertionError naming the first stage outside its tolerance."""
    cfg = WORKLOADS[workload]
    assert cfg["model"]["task"] == "flow" and cfg["model"]["reg_refine"]
    sd = synthetic_state_dict(seed=326, **(weights or BENCH_WEIGHTS), **cfg["model"])
    batch = synthetic_batch("flow", 1, H, W)
    call = cfg["call"]
    taps = {}
    mk = {k: cfg["model"][k] for k in ("num_scales", "upsample_factor", "reg_refine")}
    ref_out = O.forward(sd, batch["img0"], batch["img1"], taps=taps, **mk, **call)["flow_preds"][-1]

    m = UniMatch(**cfg["model"]).eval()
    m.load_state_dict(sd, strict=True)
    m = m.to(dev)
    ops = torch.ops.unimatch_sm100
    res = {}

    def check_feat(name, got, ref, tol):
        e = feat_err(got, ref)
        res[name] = e
        report("%-22s rel max err %.3e (tol %.1e)" % (name, e, tol))
        assert e <= tol, "%s: %.3e > %.1e" % (name, e, tol)

    def check_flow(name, got, ref, mean_tol=FLOW_MEAN, max_tol=FLOW_MAX):
        mean, mx = flow_err(got, ref)
        res[name] = (mean, mx)
        report("%-22s EPE mean %.3e max %.3e px (tol %.1e / %.1e)" % (name, mean, mx, mean_tol, max_tol))
        assert mean <= mean_tol and mx <= max_tol, "%s: mean %.3e max %.3e" % (name, mean, mx)

    with torch.no_grad():
        P = m._prepared()
        img0, img1 = batch["img0"].to(dev), batch["img1"].to(dev)
        # ---- encoder (backbone.py:104-133, trident_conv.py:64-70)
        feats = m._stage_backbone(P, img0, img1, True)
        for s, f in enumerate(feats):
            check_feat("s%d.encoder.view0" % s, f[:1], taps["s%d.f0_ori" % s], FEAT_TOL)
            check_feat("s%d.encoder.view1" % s, f[1:], taps["s%d.f1_ori" % s], FEAT_TOL)
        flow_prev = None
        for s in range(mk["num_scales"]):
            f0_ori, f1_ori = cl(taps["s%d.f0_ori" % s], dev), cl(taps["s%d.f1_ori" % s], dev)
            _, h, wd, c = f0_ori.shape
            splits, radius, prop_r = call["attn_splits_list"][s], call["corr_radius_list"][s], call["prop_radius_list"][s]
            flow_up = None
            if s > 0:
                # ---- x2 upsampling + warp (unimatch.py:154-168, geometry.py:65-72)
                flow_up = ops.upsample2x(cl(flow_prev, dev), 2.0)
                warped = ops.flow_warp(f1_ori, flow_up, h, wd)
                check_feat("s%d.warp" % s, warped, taps["s%d.f1_in" % s], FEAT_TOL)
            # ---- position + transformer (utils.py:111-131, transformer.py:226-294) on the oracle's inputs
            tok = m._stage_features(cl(taps["s%d.f0_in" % s], dev), cl(taps["s%d.f1_in" % s], dev), None, h, wd, splits)
            tok_out, _ = m._stage_transformer(P, tok, h, wd, call["attn_type"], splits, "s%d" % s)
            check_feat("s%d.transformer.view0" % s, tok_out[:1], taps["s%d.f0_tr" % s], TRANSFORMER_TOL)
            check_feat("s%d.transformer.view1" % s, tok_out[1:], taps["s%d.f1_tr" % s], TRANSFORMER_TOL)
            # ---- correlation + softmax (matching.py:7-83) on the oracle's transformer outputs
            tok_ref = torch.cat((cl(taps["s%d.f0_tr" % s], dev), cl(taps["s%d.f1_tr" % s], dev)), 0).view(2, h * wd, c)
            pred = m._stage_correlation(tok_ref, 1, h, wd, "flow", radius)
            flow = pred if flow_up is None else flow_up + pred
            check_flow("s%d.correlation" % s, flow, taps["s%d.flow_corr" % s])
            # ---- propagation (attention.py:184-253) on the oracle's features and flow
            rows = 2 * h * wd
            x_s = torch.zeros((2, (rows + 15) // 16 * 16, c), device=dev, dtype=torch.float16)
            ops.split_planes(tok_ref.view(rows, c), x_s, 0)
            flow = m._stage_propagation(P, x_s, cl(taps["s%d.flow_corr" % s], dev), 1, h, wd, prop_r)
            check_flow("s%d.propagation" % s, flow, taps["s%d.flow_prop" % s])
            flow_prev = taps["s%d.flow_prop" % s]
        # ---- refinement iterations (unimatch.py:272-354, reg_refine.py:106-119), each from the oracle's previous flow
        s = mk["num_scales"] - 1
        g0, g1 = cl(taps["s%d.f0_ori" % s], dev), cl(taps["s%d.f1_ori" % s], dev)
        feat0 = cl(taps["s%d.f0_tr" % s], dev)
        rst = m._stage_refine_setup(P, feat0, 1, h, wd)
        n_it = call["num_reg_refine"]
        mask = None
        for it in range(n_it):
            fin = cl(taps["s%d.flow_prop" % s] if it == 0 else taps["refine%d.flow" % (it - 1)], dev)
            fout, mask = m._stage_refine_iter(P, rst, g0, g1, fin, "flow", it == n_it - 1)
            check_flow("refine%d" % it, fout, taps["refine%d.flow" % it])
        check_feat("refine.mask", mask, taps["refine%d.mask" % (n_it - 1)], 1e-4)
        # ---- convex upsampling (utils.py:134-152) of the oracle's final flow with the oracle's mask
        F = mk["upsample_factor"]
        up = ops.convex_upsample(cl(taps["refine%d.flow" % (n_it - 1)], dev), cl(taps["refine%d.mask" % (n_it - 1)], dev), F, float(F))
        d = (up.cpu() - ref_out).norm(dim=1)
        res["convex_upsample"] = (d.mean().item(), d.max().item())
        report("%-22s EPE mean %.3e max %.3e px" % ("convex_upsample", d.mean().item(), d.max().item()))
        assert d.mean().item() <= FLOW_MEAN * F and d.max().item() <= FLOW_MAX * F
        # ---- and free-running end to end
        out = m(img0, img1, **call)["flow_preds"][-1]
        d = (out.cpu() - ref_out).norm(dim=1)
        res["e2e"] = (d.mean().item(), d.max().item())
        report("%-22s EPE mean %.3e max %.3e px (tol %.1e / %.1e), mean |flow| %.2f px"
               % ("end to end", d.mean().item(), d.max().item(), E2E_MEAN, E2E_MAX, ref_out.norm(dim=1).mean().item()))
        assert d.mean().item() <= E2E_MEAN and d.max().item() <= E2E_MAX
    return res

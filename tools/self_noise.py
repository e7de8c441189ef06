"""Self-noise of the REFERENCE algorithm (oracle == reference bit-for-bit) at a BASELINE shape: how far does the
reference's own fp32 output move under perturbations that are below fp32 resolution for any other correct fp32
implementation?  Two probes on the same weights and inputs:

  * `ulp`     : inputs scaled by (1 + 1e-7) (about one fp32 ulp)
  * `threads` : the same code on a different number of CPU threads (different summation orders in matmul / conv)

The end-to-end tolerance of a workload is stated as a multiple of this floor (tests/golden/cases.py); the bench weight
set is chosen so that the floor at 480x832 is far below a pixel (tools/self_noise.py --sweep prints the candidates).

    python tools/self_noise.py [--workload gmflow-scale2-regrefine6] [--size 480 832] [--damp 0.5] [--refine-gain 0.02]
                               [--backbone-gain 1.0] [--norm-gain 1.0] [--mask-gain 1.0] [--bench-set] [--threads 8 4] [--sweep]
"""
import argparse
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import unimatch_oracle as O  # noqa: E402
from unimatch_b200.spec import WORKLOADS  # noqa: E402
from unimatch_b200.synthetic import synthetic_batch, synthetic_state_dict  # noqa: E402


def run(sd, cfg, batch, threads, scale=1.0):
    torch.set_num_threads(threads)
    mk = {k: cfg["model"][k] for k in ("num_scales", "upsample_factor", "reg_refine")}
    t0 = time.perf_counter()
    out = O.forward(sd, batch["img0"] * scale, batch["img1"] * scale, intrinsics=batch.get("intrinsics"),
                    pose=batch.get("pose"), **mk, **cfg["call"])["flow_preds"][-1]
    return out, time.perf_counter() - t0


def err(a, b):
    d = (a - b).norm(dim=1) if a.dim() == 4 else (a - b).abs()
    return d.mean().item(), d.max().item()


def measure(workload, h, w, weights, threads, pair_index=0):
    cfg = WORKLOADS[workload]
    sd = synthetic_state_dict(seed=326, **weights, **cfg["model"])
    batch = synthetic_batch(cfg["model"]["task"], 1, h, w, first_index=pair_index)
    base, sec = run(sd, cfg, batch, threads[0])
    res = {"workload": workload, "size": [h, w], "weights": weights, "sec_per_forward": round(sec, 1),
           "mean_abs_output": base.abs().mean().item()}
    pert, _ = run(sd, cfg, batch, threads[0], scale=1.0 + 1e-7)
    res["ulp"] = dict(zip(("mean", "max"), err(base, pert)))
    if len(threads) > 1:
        other, _ = run(sd, cfg, batch, threads[1])
        res["threads"] = dict(zip(("mean", "max"), err(base, other)), counts=threads)
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="gmflow-scale2-regrefine6")
    ap.add_argument("--size", type=int, nargs=2, default=[480, 832])
    ap.add_argument("--damp", type=float, default=0.5)
    ap.add_argument("--refine-gain", type=float, default=0.02)
    ap.add_argument("--backbone-gain", type=float, default=1.0)
    ap.add_argument("--norm-gain", type=float, default=1.0)
    ap.add_argument("--mask-gain", type=float, default=1.0)
    ap.add_argument("--bench-set", action="store_true", help="use synthetic.BENCH_WEIGHTS (the well-conditioned set bench.py loads)")
    ap.add_argument("--threads", type=int, nargs="+", default=[8, 4])
    ap.add_argument("--sweep", action="store_true")
    args = ap.parse_args()
    sets = [dict(damp=args.damp, refine_gain=args.refine_gain, backbone_gain=args.backbone_gain, norm_gain=args.norm_gain,
                 mask_gain=args.mask_gain)]
    if args.bench_set:
        from unimatch_b200.synthetic import BENCH_WEIGHTS
        sets = [dict(BENCH_WEIGHTS)]
    if args.sweep:
        sets = [dict(damp=d, refine_gain=g) for d in (0.5, 0.35, 0.25) for g in (0.02, 0.005)]
    for ws in sets:
        print(json.dumps(measure(args.workload, args.size[0], args.size[1], ws, args.threads)), flush=True)


if __name__ == "__main__":
    main()

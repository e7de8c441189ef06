"""A/B harness (GPU box): the drop-in module against the library path, i.e. the oracle port -- the reference's own eager
ATen op sequence (cuBLAS / cuDNN / ATen kernels) -- run on the same GPU, with TF32 off (fp32-exact libraries) and on (the
torch default for convolutions).  Prints time per forward and the difference to the fp32 CPU oracle for each, so a
regression can be attributed to our kernels or to the libraries' arithmetic.  Nothing here is on the product path.

    python tools/ab_paths.py [--workload gmflow-scale2-regrefine6] [--size 480 832] [--batch 2]
"""
import argparse
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import unimatch_oracle as O  # noqa: E402
from unimatch_b200 import UniMatch  # noqa: E402
from unimatch_b200.spec import WORKLOADS  # noqa: E402
from unimatch_b200.synthetic import BENCH_WEIGHTS, synthetic_batch, synthetic_state_dict  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="gmflow-scale2-regrefine6")
    ap.add_argument("--size", type=int, nargs=2, default=[480, 832])
    ap.add_argument("--batch", type=int, default=2)
    a = ap.parse_args()
    cfg = WORKLOADS[a.workload]
    sd = synthetic_state_dict(seed=326, **BENCH_WEIGHTS, **cfg["model"])
    data = synthetic_batch(cfg["model"]["task"], a.batch, *a.size)
    mk = {k: cfg["model"][k] for k in ("num_scales", "upsample_factor", "reg_refine")}

    def oracle(dev):
        d = {k: v.to(dev) for k, v in data.items()}
        s = {k: v.to(dev) for k, v in sd.items()}
        return O.forward(s, d["img0"], d["img1"], intrinsics=d.get("intrinsics"), pose=d.get("pose"), **mk, **cfg["call"])["flow_preds"][-1]

    ref = oracle("cpu")
    err = lambda x: ((x.cpu() - ref).norm(dim=1) if ref.dim() == 4 else (x.cpu() - ref).abs())
    m = UniMatch(**cfg["model"]).eval()
    m.load_state_dict(sd)
    m = m.cuda()
    dev = {k: v.cuda() for k, v in data.items()}
    arms = {"libunimatch_sm100 (ours)": lambda: m(dev["img0"], dev["img1"], intrinsics=dev.get("intrinsics"), pose=dev.get("pose"),
                                                   **cfg["call"])["flow_preds"][-1]}
    for tf32 in (False, True):
        def lib(tf32=tf32):
            torch.backends.cuda.matmul.allow_tf32 = tf32
            torch.backends.cudnn.allow_tf32 = tf32
            return oracle("cuda")
        arms["reference eager on cuda, TF32 %s" % ("on" if tf32 else "off")] = lib
    for name, fn in arms.items():
        fn(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        out = fn(); torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        e = err(out)
        print("%-40s %8.1f ms/forward (%d pairs)   vs fp32 CPU oracle: mean %.3e max %.3e" % (name, dt * 1e3, a.batch, e.mean().item(), e.max().item()))


if __name__ == "__main__":
    main()

"""Drop-in proof (SURVEY.md section 8b "Integration proof"): run the reference's UNMODIFIED `main_flow.main()` inference
entry (/root/reference/main_flow.py:136-366 -> evaluate_flow.inference_flow) twice on the reference's own demo frames --
once as shipped, once with `main_flow.UniMatch` (bound at main_flow.py:10) replaced by `unimatch_b200.UniMatch` -- loading
the same checkpoint through the reference's own `--resume` path, and compare the `.flo` files it writes.

Runs on the CPU of the build container (the only place /root/reference exists): the module's custom ops get the
oracle-backed CPU kernels of tests/refops.py, so what is proven is the boundary -- constructor, state_dict, forward
signature, output dict, dtype/shape/layout conventions, eval()/no_grad usage -- not the CUDA kernels (tests -m gpu).
`imageio`, `skimage(.io)` and `matplotlib(.cm)` are import-time-only dependencies of the reference's drivers that this
image lacks; they are stubbed.  Usage: python tests/dropin_main_flow.py OUTDIR  -> prints one JSON line."""
import json
import os
import sys
import types

import numpy as np
import torch

REF = "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.dont_write_bytecode = True          # the reference mount is read-only
for p in (REF, ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)

for name in ("imageio", "skimage", "skimage.io", "matplotlib", "matplotlib.cm", "matplotlib.pyplot"):
    if name not in sys.modules:
        try:
            __import__(name)
        except Exception:
            sys.modules[name] = types.ModuleType(name)
if not hasattr(sys.modules["skimage"], "io"):
    sys.modules["skimage"].io = sys.modules["skimage.io"]
if not hasattr(sys.modules["matplotlib"], "cm"):
    sys.modules["matplotlib"].cm = sys.modules["matplotlib.cm"]
if not hasattr(sys.modules["matplotlib.cm"], "get_cmap"):          # imported by name (utils/visualization.py:6), never called here
    sys.modules["matplotlib.cm"].get_cmap = lambda *a, **k: None


def read_flo(path):
    with open(path, "rb") as f:
        assert np.fromfile(f, np.float32, 1)[0] == 202021.25
        w, h = np.fromfile(f, np.int32, 2)
        return np.fromfile(f, np.float32, 2 * w * h).reshape(h, w, 2)


def main():
    out = sys.argv[1]
    torch.set_num_threads(8)
    import main_flow                                     # the reference entry script, unmodified
    from unimatch_b200.spec import WORKLOADS
    from unimatch_b200.synthetic import BENCH_WEIGHTS, synthetic_state_dict
    wl = WORKLOADS["gmflow-scale1"]
    ckpt = os.path.join(out, "synthetic.pth")
    torch.save({"model": synthetic_state_dict(seed=326, **BENCH_WEIGHTS, **wl["model"])}, ckpt)

    def run(tag):
        d = os.path.join(out, tag)
        argv = ["--inference_dir", os.path.join(REF, "demo", "flow-davis"), "--output_path", d, "--inference_size", "256", "448",
                "--resume", ckpt, "--strict_resume", "--save_flo_flow", "--padding_factor", "16", "--upsample_factor", "8",
                "--num_scales", "1", "--attn_splits_list", "2", "--corr_radius_list", "-1", "--prop_radius_list", "-1"]
        main_flow.main(main_flow.get_args_parser().parse_args(argv))
        return {f: read_flo(os.path.join(d, f)) for f in sorted(os.listdir(d)) if f.endswith(".flo")}

    ref = run("reference")
    import refops
    import unimatch_b200
    refops.register_cpu_kernels()
    main_flow.UniMatch = unimatch_b200.UniMatch          # the drop-in: nothing else changes
    ours = run("dropin")
    assert sorted(ref) == sorted(ours) and ref, (sorted(ref), sorted(ours))
    res = {"files": sorted(ref), "shape": list(next(iter(ref.values())).shape)}
    epe = [np.linalg.norm(ref[f] - ours[f], axis=-1) for f in ref]
    res["mean_epe"] = float(np.mean([e.mean() for e in epe]))
    res["max_epe"] = float(np.max([e.max() for e in epe]))
    res["mean_flow"] = float(np.mean([np.linalg.norm(v, axis=-1).mean() for v in ref.values()]))
    print("DROPIN " + json.dumps(res))


if __name__ == "__main__":
    main()

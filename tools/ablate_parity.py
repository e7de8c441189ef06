"""Which tensor-core component contributes how much end-to-end error?  (GPU box)"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import cases
from unimatch_b200 import UniMatch, ops
gold = torch.load(os.path.join(ROOT, "tests", "golden", "golden.pt"))["vectors"]
names = sys.argv[1:] or ["e2e_gmstereo_s2", "e2e_gmstereo_s2_rr3", "e2e_gmflow_s2_rr6", "e2e_gmflow_s1_256x320"]
for name in names:
    cfg, sd, batch, call = cases.e2e_setup(name)
    m = UniMatch(**cfg["model"]).eval(); m.load_state_dict(sd); m = m.cuda()
    dev = {k: v.cuda() for k, v in batch.items()}
    for label, (bb, gm, cv, simt) in {"all tc": (1, 1, 1, 0), "no tc backbone": (0, 1, 1, 0), "no tc gemm": (1, 0, 1, 0), "no tc conv": (1, 1, 0, 0),
                                      "attention on cuda cores": (1, 1, 1, 1), "nothing on tc": (0, 0, 0, 1)}.items():
        m.tc_backbone, m.tc_gemm, m.tc_conv = bool(bb), bool(gm), bool(cv)
        ops.set_force_cuda_cores(bool(simt))
        out = m(dev["img0"], dev["img1"], intrinsics=dev.get("intrinsics"), pose=dev.get("pose"), **call)["flow_preds"][-1]
        mean, mx = cases.epe(out.cpu(), gold[name])
        print("%-24s %-26s mean %.3e max %.3e (x%.1f thread-noise)" % (name, label, mean, mx, mean / cases.E2E_NOISE[name]), flush=True)
    ops.set_force_cuda_cores(False)

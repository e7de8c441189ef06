"""The fused attention kernel alone on window-major operand planes at the bench shapes (8 pairs):
   ncu --set full --clock-control none --import-source on -k regex:attn_tc -s 2 -c 2 -o gpurun_out/prof_attn python tools/profile_attn.py
   python tools/profile_attn.py --time      (CUDA events, 20 launches per class)"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from unimatch_b200 import ops  # noqa: E402

OPS = torch.ops.unimatch_sm100
torch.manual_seed(0)
n = 16


def case(h, w, K, shift):
    lp = ops.attention_planes_lp(h, w, K, K, 0, 0, 0)
    lw = (h // K) * (w // K)
    pl = []
    for _ in range(3):
        t = torch.zeros((2, n, K * K, lp, 128), device="cuda", dtype=torch.float16)
        t[:, :, :, :lw] = (torch.randn((2, n, K * K, lw, 128), device="cuda") * (1.5 if len(pl) < 2 else 1.0)).half()
        t[1] *= 1e-3
        pl.append(t)
    out_s = torch.empty((2, n * h * w, 128), device="cuda", dtype=torch.float16)
    sh, sw = ((h // K) // 2, (w // K) // 2) if shift else (0, 0)
    mask = ops.MASK_SWIN if shift else ops.MASK_NONE
    fl = 4.0 * lw * lw * 128 * K * K * n
    return (lambda: OPS.window_attention_planes(pl[0], pl[1], pl[2], n, n // 2, h, w, K, K, sh, sw, mask, None, out_s)), fl


named = [("s0 60x104 K=2 shifted", *case(60, 104, 2, True)), ("s1 120x208 K=8 shifted", *case(120, 208, 8, True)),
         ("s0 60x104 K=2 plain", *case(60, 104, 2, False)), ("s1 120x208 K=8 plain", *case(120, 208, 8, False))]
if "--time" in sys.argv:
    for name, f, fl in named:
        for _ in range(3):
            f()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        t0.record()
        for _ in range(20):
            f()
        t1.record()
        torch.cuda.synchronize()
        ms = t0.elapsed_time(t1) / 20
        print("%-26s %.4f ms  %.1f TFLOP/s algorithmic" % (name, ms, fl / ms / 1e9), flush=True)
else:
    for rep in range(2):
        for _, f, _ in named[:2]:
            f()
        torch.cuda.synchronize()
print("done")

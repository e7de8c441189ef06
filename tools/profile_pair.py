"""The CTA-pair kernels of round 2b at the bench shapes (8 pairs, 120x208 features / 399 360 token rows), two launches each:
   ncu --metrics <list in tools/gpu_round.sh> --clock-control none -k regex:"conv_tc_kernel|ffn_tc_kernel" -s 4 -c 4 ...
   GRU z|r 1x5 as two 128-wide pair tiles, convc2 3x3 256 -> 192 as two 96-wide pair tiles, the fused FFN at scale 1 and 0."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from unimatch_b200 import ops  # noqa: E402
from tools.profile_kernels import ffn_fused  # noqa: E402,F401  (also runs nothing: guarded below)

OPS = torch.ops.unimatch_sm100
torch.manual_seed(0)
B, h, w = 8, 120, 208


def gru_zr128():
    x_s = torch.randn((2, B, h, w, 128), device="cuda").half()
    wt = ops.prep_conv_weight(torch.randn(256, 128, 1, 5, device="cuda") * 0.03, [128], 256)
    hh = torch.randn((B, h, w, 128), device="cuda")
    pre = torch.randn((B, h, w, 256), device="cuda")
    z = torch.empty((B, h, w, 128), device="cuda")
    rh_s = torch.empty((2, B, h, w, 128), device="cuda", dtype=torch.float16)
    return lambda: OPS.conv2d_tc(x_s, None, wt, None, 1, 5, 0, 2, 256, 128, ops.CONV_GRU_ZR, 0, z, 0, rh_s, 0, hh, None, pre=pre)


def convc2_96():
    x_s = torch.randn((2, B, h, w, 256), device="cuda").half()
    wt = ops.prep_conv_weight(torch.randn(192, 256, 3, 3, device="cuda") * 0.02, [256], 192)
    bias = torch.zeros(192, device="cuda")
    o_s = torch.empty((2, B, h, w, 256), device="cuda", dtype=torch.float16)
    return lambda: OPS.conv2d_tc(x_s, None, wt, bias, 3, 3, 1, 1, 192, 96, ops.CONV_LINEAR, ops.ACT_RELU, None, 0, o_s, 0, None, None)


rows1 = 2 * B * h * w
named = [("gru z|r 1x5 128->256 (hoisted K = 640), 2 x 128-wide pair tiles", gru_zr128()),
         ("convc2 3x3 256->192, 2 x 96-wide pair tiles", convc2_96()),
         ("fused FFN scale 1 (399 360 rows)", ffn_fused(rows1)), ("fused FFN scale 0 (99 840 rows)", ffn_fused(rows1 // 4))]
if "--time" in sys.argv:
    for name, f in named:
        for _ in range(3):
            f()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        t0.record()
        for _ in range(10):
            f()
        t1.record()
        torch.cuda.synchronize()
        print("%-70s %.3f ms" % (name, t0.elapsed_time(t1) / 10), flush=True)
else:
    for rep in range(2):
        for _, f in named:
            f()
        torch.cuda.synchronize()
print("done")

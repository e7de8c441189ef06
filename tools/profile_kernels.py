"""Launch each hot tensor-core kernel twice (warm-up, then the launch ncu captures):
   ncu --set full --clock-control none --import-source on -k regex:"attn_tc_kernel|conv_tc_kernel" -s 4 -c 4 \
       -o gpurun_out/prof python tools/profile_kernels.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from unimatch_b200 import ops  # noqa: E402

OPS = torch.ops.unimatch_sm100
torch.manual_seed(0)
pairs = 8
n = 2 * pairs


def attn(h, w, K):
    d = torch.randn((n, h * w, 384), device="cuda")
    wh, ww = h // K, w // K
    return lambda: OPS.window_attention(d[..., :128], d[..., 128:256], d[..., 256:], pairs, h, w, K, K, wh // 2, ww // 2,
                                        ops.MASK_SWIN)


def gemm_in(rows):
    x_s = torch.randn((2, 1, rows // 16, 16, 128), device="cuda").half()
    wt = ops.prep_conv_weight(torch.randn(640, 128, 1, 1, device="cuda") * 0.1, [128], 640)
    y = torch.empty((1, rows // 16, 16, 640), device="cuda")
    return lambda: OPS.conv2d_tc(x_s, None, wt, None, 1, 1, 0, 0, 640, 128, ops.CONV_LINEAR, ops.ACT_NONE, y, 0, None, 0,
                                 None, None)


def gru_zr(b, h, w):
    h_s = torch.randn((2, b, h, w, 128), device="cuda").half()
    x_s = torch.randn((2, b, h, w, 256), device="cuda").half()
    wt = ops.prep_conv_weight(torch.randn(256, 384, 1, 5, device="cuda") * 0.03, [128, 256], 256)
    bias = torch.zeros(256, device="cuda")
    hh = torch.randn((b, h, w, 128), device="cuda")
    z = torch.empty((b, h, w, 128), device="cuda")
    rh_s = torch.empty((2, b, h, w, 128), device="cuda", dtype=torch.float16)
    return lambda: OPS.conv2d_tc(h_s, x_s, wt, bias, 1, 5, 0, 2, 256, 128, ops.CONV_GRU_ZR, 0, z, 0, rh_s, 0, hh, None)


fns = [attn(60, 104, 2), attn(120, 208, 8), gemm_in(n * 120 * 208), gru_zr(pairs, 120, 208)]
for rep in range(2):
    for f in fns:
        f()
    torch.cuda.synchronize()
print("done")

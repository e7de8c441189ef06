"""Launch each hot tensor-core kernel twice (warm-up, then the launch ncu captures):
   ncu --set full --clock-control none --import-source on -k regex:"attn_tc_kernel|conv_tc_kernel" -s 7 -c 7 \
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
    return lambda: OPS.conv2d_tc(h_s, x_s, wt, bias, 1, 5, 0, 2, 256, 256, ops.CONV_GRU_ZR, 0, z, 0, rh_s, 0, hh, None)


def ffn1(rows):
    a_s = torch.randn((2, 1, rows // 16, 16, 128), device="cuda").half()
    b_s = torch.randn((2, 1, rows // 16, 16, 128), device="cuda").half()
    wt = ops.prep_conv_weight(torch.randn(1024, 256, 1, 1, device="cuda") * 0.1, [128, 128], 1024)
    hid = torch.empty((2, 1, rows // 16, 16, 1024), device="cuda", dtype=torch.float16)
    bn = 128 if '--ffn1-128' in sys.argv else 256
    return lambda: OPS.conv2d_tc(a_s, b_s, wt, None, 1, 1, 0, 0, 1024, bn, ops.CONV_LINEAR, ops.ACT_GELU, None, 0, hid, 0,
                                 None, None)


def ffn2(rows):
    hid = torch.randn((2, 1, rows // 16, 16, 1024), device="cuda").half()
    wt = ops.prep_conv_weight(torch.randn(128, 1024, 1, 1, device="cuda") * 0.05, [1024], 128)
    res = torch.randn((1, rows // 16, 16, 128), device="cuda")
    o_f = torch.empty((1, rows // 16, 16, 128), device="cuda")
    o_s = torch.empty((2, 1, rows // 16, 16, 128), device="cuda", dtype=torch.float16)
    gam, bet = torch.ones(128, device="cuda"), torch.zeros(128, device="cuda")
    return lambda: OPS.conv2d_tc(hid, None, wt, None, 1, 1, 0, 0, 128, 128, ops.CONV_LN, 0, o_f, 0, o_s, 0, res, None, gam, bet)


def merge_ln(rows):
    m_s = torch.randn((2, 1, rows // 16, 16, 128), device="cuda").half()
    wt = ops.prep_conv_weight(torch.randn(128, 128, 1, 1, device="cuda") * 0.1, [128], 128)
    res = torch.randn((1, rows // 16, 16, 128), device="cuda")
    o_f = torch.empty((1, rows // 16, 16, 128), device="cuda")
    o_s = torch.empty((2, 1, rows // 16, 16, 128), device="cuda", dtype=torch.float16)
    gam, bet = torch.ones(128, device="cuda"), torch.zeros(128, device="cuda")
    return lambda: OPS.conv2d_tc(m_s, None, wt, None, 1, 1, 0, 0, 128, 128, ops.CONV_LN, 0, o_f, 0, o_s, 0, res, None, gam, bet)


def ffn_fused(rows):
    a_s = torch.randn((2, rows, 128), device="cuda").half()
    b_s = torch.randn((2, rows, 128), device="cuda").half()
    w1 = ops.prep_conv_weight(torch.randn(1024, 256, 1, 1, device="cuda") * 0.1, [128, 128], 1024)
    w2 = ops.prep_conv_weight(torch.randn(128, 1024, 1, 1, device="cuda") * 0.05, [1024], 128)
    res = torch.randn((rows, 128), device="cuda")
    o_f = torch.empty((rows, 128), device="cuda")
    o_s = torch.empty((2, rows, 128), device="cuda", dtype=torch.float16)
    gam, bet = torch.ones(128, device="cuda"), torch.zeros(128, device="cuda")
    return lambda: OPS.ffn_tc(a_s, b_s, w1, w2, res, gam, bet, o_f, o_s, rows)


if __name__ == "__main__":
    rows1 = n * 120 * 208
    named = [("attn s0 (60x104, K=2)", attn(60, 104, 2)), ("attn s1 (120x208, K=8)", attn(120, 208, 8)),
             ("gemm_in 128->640 s1", gemm_in(rows1)), ("gru z|r 1x5 s1", gru_zr(pairs, 120, 208)),
             ("ffn1 256->1024 gelu s1", ffn1(rows1)), ("ffn2 1024->128 ln s1", ffn2(rows1)), ("merge 128->128 ln s1", merge_ln(rows1)),
             ("ffn fused 256->1024->128 s1", ffn_fused(rows1)), ("ffn fused s0", ffn_fused(rows1 // 4))]
    if "--time" in sys.argv:                                   # CUDA-event timing of each launch class (no profiler)
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
            print("%-28s %.3f ms" % (name, t0.elapsed_time(t1) / 10), flush=True)
    else:
        for rep in range(2):
            for _, f in named:
                f()
            torch.cuda.synchronize()
    print("done")

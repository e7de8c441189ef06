"""One launch of every HBM / L2-bound kernel of the path at the bench shapes (8 pairs, 480x832 -> 120x208 features), for
   ncu --set full --clock-control none --import-source on -o gpurun_out/prof_hbm python tools/profile_hbm.py
and, with --time, CUDA-event timings (10 launches each, inputs >> L2) with the algorithmic bytes and the achieved GB/s."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from unimatch_b200 import ops  # noqa: E402

OPS = torch.ops.unimatch_sm100
torch.manual_seed(0)
B, h, w, C = 8, 120, 208, 128
dev = "cuda"
f0 = torch.randn((B, h, w, C), device=dev)
f1 = torch.randn((B, h, w, C), device=dev)
flow = torch.randn((B, h, w, 2), device=dev) * 3          # incoherent: neighbouring pixels point anywhere
smooth = torch.nn.functional.interpolate(torch.randn((B, 2, 8, 13), device=dev) * 8, size=(h, w), mode="bilinear",
                                         align_corners=True).permute(0, 2, 3, 1).contiguous()   # a smooth field, like a real flow
mask = torch.randn((B, h, w, 144), device=dev)
img0 = torch.rand((B, 3, 480, 832), device=dev) * 255
img1 = torch.rand((B, 3, 480, 832), device=dev) * 255
stem_w = torch.randn((64, 3, 7, 7), device=dev) * 0.1
stem_out = torch.empty((2 * B, 240, 416, 64), device=dev)
a64 = torch.randn((2 * B, 240, 416, 64), device=dev)
pl64 = torch.empty((2, 2 * B, 240, 416, 64), device=dev, dtype=torch.float16)
o64 = torch.empty_like(a64)
corr = torch.randn((B, h, w, 81), device=dev)
pl128 = torch.zeros((2, B, h, w, 128), device=dev, dtype=torch.float16)
qk = torch.randn((B, h * w, 256), device=dev)
full = torch.randn((B, 2, 480, 832), device=dev)
tok = torch.randn((2 * B, h, w, C), device=dev)
table = torch.randn((15, 26, C), device=dev)
convf1_w = torch.randn((128, 2, 7, 7), device=dev) * 0.1
convf1_b = torch.zeros(128, device=dev)
flo_s = torch.empty((2, B, h, w, 128), device=dev, dtype=torch.float16)
st64 = OPS.instance_norm_stats(a64)
MB = 1e6
px = B * h * w
named = [
    # name, callable, algorithmic bytes (reads + writes every byte once)
    ("flow_warp", lambda: OPS.flow_warp(f1, flow, h, w), px * (512 + 8 + 512)),
    ("local_corr_softmax r4", lambda: OPS.local_corr_softmax(f0, f1, h, w, 4, 4, False), px * (512 + 512 + 8)),
    ("local_corr_volume r4 (smooth flow)", lambda: OPS.local_corr_volume(f0, f1, smooth, h, w, 4), px * (512 + 512 + 8 + 324)),
    ("local_corr_volume r4 (noise flow)", lambda: OPS.local_corr_volume(f0, f1, flow, h, w, 4), px * (512 + 512 + 8 + 324)),
    ("propagate_local r1", lambda: OPS.propagate_local(qk[:, :, :128], qk[:, :, 128:], flow, h, w, 1), px * (512 + 512 + 8 + 8)),
    ("convex_upsample x4", lambda: OPS.convex_upsample(flow, mask, 4, 4.0), px * (576 + 8 + 16 * 8)),
    ("instance_norm_stats 64ch", lambda: OPS.instance_norm_stats(a64), a64.numel() * 4),
    ("instance_norm_apply 64ch", lambda: OPS.instance_norm_apply(a64, st64, True, None, None, False, o64, pl64, 0), a64.numel() * (4 + 4 + 4)),
    ("split_planes 81ch", lambda: OPS.split_planes(corr, pl128, 0), px * 81 * 8),
    ("conv7x7 stem (3->64, s2)", lambda: OPS.conv7x7_small(img0, img1, True, stem_w, None, 2, False, [1.0, 1.0, 1.0], [0.0, 0.0, 0.0], stem_out, None),
     2 * B * 3 * 480 * 832 * 4 + stem_out.numel() * 4),
    ("conv7x7 flow encoder (2->128)", lambda: OPS.conv7x7_small(flow, None, False, convf1_w, convf1_b, 1, True, None, None, None, flo_s), px * (8 + 512)),
    ("add_position", lambda: OPS.add_position(tok, table, h, w), tok.numel() * 8),
    ("upsample2x", lambda: OPS.upsample2x(flow[:, :60, :104].contiguous(), 2.0), px * 8 * 1.25),
    ("resize_bilinear 480x832", lambda: OPS.resize_bilinear(full, 436, 1024, [1.2, 0.9], False), full.numel() * 4 + B * 2 * 436 * 1024 * 4),
    ("fb_consistency 480x832", lambda: OPS.fb_consistency(full[:4], full[4:], 0.01, 0.5), 4 * 480 * 832 * (16 + 8)),
]
if "--time" in sys.argv:
    peak = 6566.4
    print("| kernel | ms | algorithmic MB | achieved GB/s | fraction of %.0f GB/s (measured copy peak) |\n|---|---|---|---|---|" % peak)
    for name, f, nbytes in named:
        for _ in range(3):
            f()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        t0.record()
        for _ in range(10):
            f()
        t1.record()
        torch.cuda.synchronize()
        ms = t0.elapsed_time(t1) / 10
        gbs = nbytes / (ms * 1e-3) / 1e9
        print("| %s | %.4f | %.1f | %.0f | %.3f |" % (name, ms, nbytes / MB, gbs, gbs / peak), flush=True)
else:
    for name, f, _ in named:
        f()
    torch.cuda.synchronize()
print("done")

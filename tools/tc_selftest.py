"""GPU bring-up harness for the tcgen05 attention kernel (run on the B200 box, each stage in its own process so a
trapped kernel does not hide the other stages):  python tools/tc_selftest.py [dump|parity|perf|all]"""
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def stage_dump():
    import ctypes
    import torch
    from unimatch_b200 import ops
    OPS = torch.ops.unimatch_sm100
    torch.manual_seed(0)
    n, h, w = 2, 16, 16                      # one window of 256 tokens: 2 query tiles, 4 key tiles
    L = h * w
    q = torch.randn(n, L, 128, device="cuda")
    k = torch.randn(n, L, 128, device="cuda")
    v = torch.randn(n, L, 128, device="cuda")
    dump = torch.zeros(128 * 64 + 128 * 128, device="cuda")
    ops.LIB.um_debug_set_dump(ctypes.c_void_p(dump.data_ptr()))
    out = OPS.window_attention(q, k, v, 0, h, w, 1, 1, 0, 0, 0)
    torch.cuda.synchronize()
    ops.LIB.um_debug_set_dump(None)
    S = dump[:128 * 64].view(128, 64)
    S_ref = (q[0, :128].double() @ k[0, :64].double().t()).float()
    err_s = (S - S_ref).abs().max().item()
    print("S tile: max|S - q k^T| = %.3e  (|S_ref| max %.2f)" % (err_s, S_ref.abs().max().item()))
    if err_s > 1e-3:
        print("  S[0,:8]    =", S[0, :8].tolist())
        print("  Sref[0,:8] =", S_ref[0, :8].tolist())
        print("  S[1,:4] =", S[1, :4].tolist(), " Sref[1,:4] =", S_ref[1, :4].tolist())
        # which (row, col) permutation? correlate
        hit = (S[:, :, None] - S_ref[0, None, :]).abs().min().item()
        print("  closest match of any S entry to S_ref[0,:]:", hit)
    logits = (q[0, :128].double() @ k[0].double().t()) / (128 ** 0.5)
    P = torch.softmax(logits, -1)
    O_ref = (P @ v[0].double()).float()
    err_o = (out[0, :128] - O_ref).abs().max().item()
    print("attention rows 0..127: max|out - ref| = %.3e" % err_o)
    O_un = dump[128 * 64:].view(128, 128)
    print("un-normalised O finite:", bool(torch.isfinite(O_un).all()), " |O_un| max %.3e" % O_un.abs().max().item())
    full_ref = torch.stack([(torch.softmax((q[i].double() @ k[i].double().t()) / (128 ** 0.5), -1) @ v[i].double()).float()
                            for i in range(n)])
    print("all rows: max|out - ref| = %.3e" % (out - full_ref).abs().max().item())


def stage_parity():
    import torch
    import refops
    from unimatch_b200 import ops
    OPS = torch.ops.unimatch_sm100
    cases = [(2, 16, 16, 1, 1, False, 0), (2, 32, 24, 2, 2, False, 1), (2, 32, 24, 2, 2, True, 1),
             (2, 30, 52, 2, 2, True, 1), (2, 60, 104, 2, 2, True, 1), (2, 60, 104, 2, 2, False, 0),
             (2, 120, 208, 8, 8, True, 1), (4, 15, 26, 1, 1, False, 2)]
    for (n, h, w, kh, kw, shift, kvs) in cases:
        g = torch.Generator().manual_seed(h * w + kh)
        L = h * w
        qkv = torch.randn((n, L, 384), generator=g) * 1.5
        wh, ww = h // kh, w // kw
        sh, sw = (wh // 2, ww // 2) if shift else (0, 0)
        mask = ops.MASK_SWIN if shift else ops.MASK_NONE
        d = qkv.cuda()
        args = (d[..., :128], d[..., 128:256], d[..., 256:], kvs, h, w, kh, kw, sh, sw, mask)
        ops.set_force_cuda_cores(True)
        ref_gpu = OPS.window_attention(*args)
        ops.set_force_cuda_cores(False)
        got = OPS.window_attention(*args)
        torch.cuda.synchronize()
        e1 = (got - ref_gpu).abs().max().item()
        msg = "n=%d %dx%d K=%dx%d shift=%d kvs=%d Lw=%d: |tc - simt| = %.3e" % (n, h, w, kh, kw, shift, kvs, wh * ww, e1)
        if L <= 4000:
            ref = refops.window_attention(qkv[..., :128], qkv[..., 128:256], qkv[..., 256:], kvs, h, w, kh, kw, sh, sw, mask)
            msg += "  |tc - oracle| = %.3e  |simt - oracle| = %.3e" % ((got.cpu() - ref).abs().max().item(),
                                                                      (ref_gpu.cpu() - ref).abs().max().item())
        print(msg, flush=True)


def stage_perf():
    import torch
    from unimatch_b200 import ops
    OPS = torch.ops.unimatch_sm100
    for (pairs, h, w, K) in [(8, 60, 104, 2), (8, 120, 208, 8)]:
        n = 2 * pairs
        L = h * w
        d = torch.randn((n, L, 384), device="cuda")
        wh, ww = h // K, w // K
        args = (d[..., :128], d[..., 128:256], d[..., 256:], pairs, h, w, K, K, wh // 2, ww // 2, ops.MASK_SWIN)
        flops = 4.0 * (wh * ww) ** 2 * 128 * K * K * n
        for force in (True, False):
            ops.set_force_cuda_cores(force)
            for _ in range(3):
                OPS.window_attention(*args)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                OPS.window_attention(*args)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 10
            print("%s %dx%d K=%d n=%d: %.3f ms/call  %.1f TFLOP/s (fp32-equivalent)" %
                  ("cuda-core" if force else "tcgen05  ", h, w, K, n, ms, flops / ms / 1e9), flush=True)
        ops.set_force_cuda_cores(False)


def stage_conv():
    """Update block: tensor-core implicit GEMM vs cuDNN fp32, real shapes (8 pairs, 120x208)."""
    import torch
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    from unimatch_b200 import UniMatch
    from unimatch_b200.synthetic import synthetic_state_dict
    kw = dict(num_scales=2, upsample_factor=4, reg_refine=True, task="flow")
    m = UniMatch(**kw).eval()
    m.load_state_dict(synthetic_state_dict(seed=326, damp=0.5, **kw))
    m = m.cuda()
    P = m._prepared()
    b, h, w = 8, 120, 208
    g = torch.Generator().manual_seed(3)
    feat0 = torch.randn((b, h, w, 128), generator=g).cuda()
    corr = (torch.randn((b, h, w, 81), generator=g) * 4).cuda()
    flow = (torch.randn((b, h, w, 2), generator=g) * 2).cuda()
    import torch.nn.functional as F
    with torch.no_grad():
        torch.backends.cudnn.allow_tf32 = False
        proj = F.linear(feat0, P["proj_w"], P["raw"]["refine_proj.bias"])
        net0, inp = torch.tanh(proj[..., :128]).contiguous(), torch.relu(proj[..., 128:])
        ref = m._update_block(P, net0, inp, corr, flow, True)
        st = m._refine_setup(P, feat0, b, h, w)
        got = m._update_block_tc(P, st, corr, flow, True)
        for name, a, bb in zip(("net", "mask", "delta"), got, ref):
            print("update block %s: max|tc - cudnn| = %.3e (max|ref| %.2f)" % (name, (a - bb).abs().max().item(), bb.abs().max().item()))
        for label, fn in (("cudnn fp32", lambda: m._update_block(P, net0, inp, corr, flow, True)),
                          ("tcgen05   ", lambda: m._update_block_tc(P, st, corr, flow, True))):
            for _ in range(2):
                fn()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5):
                fn()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 5
            print("%s update block (8 pairs): %.2f ms  -> %.1f TFLOP/s (147 GFLOP/pair/iter)" % (label, ms, 147.0 * 8 / ms), flush=True)


def stage_backbone():
    import torch
    from unimatch_b200 import UniMatch
    from unimatch_b200.synthetic import synthetic_state_dict, synthetic_batch
    kw = dict(num_scales=2, upsample_factor=4, reg_refine=True, task="flow")
    m = UniMatch(**kw).eval()
    m.load_state_dict(synthetic_state_dict(seed=326, damp=0.5, **kw))
    m = m.cuda()
    P = m._prepared()
    b = synthetic_batch("flow", 4, 480, 832)
    x = torch.cat((b["img0"], b["img1"]), 0).cuda()
    mean = torch.tensor([0.485, 0.456, 0.406], device="cuda").view(1, 3, 1, 1)
    std = torch.tensor([0.229, 0.224, 0.225], device="cuda").view(1, 3, 1, 1)
    x = (x / 255.0 - mean) / std
    with torch.no_grad():
        torch.backends.cudnn.allow_tf32 = False
        ref = m._backbone(P["raw"], x)
        raw = (b["img0"].cuda(), b["img1"].cuda(), True)
        got = m._backbone_tc(P, raw)
        for a, r in zip(got, ref):
            print("backbone feature %s: max|tc - cudnn| = %.3e (max|ref| %.2f)" % (tuple(a.shape), (a - r).abs().max().item(), r.abs().max().item()))
        for label, fn in (("cudnn fp32", lambda: m._backbone(P["raw"], x)), ("tcgen05   ", lambda: m._backbone_tc(P, raw))):
            for _ in range(2):
                fn()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5):
                fn()
            e1.record()
            torch.cuda.synchronize()
            print("%s backbone (4 pairs): %.2f ms" % (label, e0.elapsed_time(e1) / 5), flush=True)


def stage_accuracy():
    """Error of the 3xFP16 tensor-core GEMM vs an fp64 reference, next to cuBLAS fp32 (SIMT) on the same data."""
    import torch
    from unimatch_b200 import ops
    OPS = torch.ops.unimatch_sm100
    torch.manual_seed(1)
    rows, K = 4096, 1152
    for N, label, mk in ((128, "signed", lambda *s: torch.randn(*s, device="cuda")), (128, "positive", lambda *s: torch.rand(*s, device="cuda")),
                         (64, "signed/bn64", lambda *s: torch.randn(*s, device="cuda")), (64, "positive/bn64", lambda *s: torch.rand(*s, device="cuda"))):
        a = mk(rows, K)
        w = mk(N, K) * (1.0 / K ** 0.5)
        a_s = torch.zeros((2, 1, rows // 16, 16, K), device="cuda", dtype=torch.float16)
        OPS.split_planes(a, a_s, 0)
        wp = ops.prep_conv_weight(w.view(N, K, 1, 1), [K], N)
        out = torch.empty((1, rows // 16, 16, N), device="cuda")
        OPS.conv2d_tc(a_s, None, wp, None, 1, 1, 0, 0, N, N, ops.CONV_LINEAR, ops.ACT_NONE, out, 0, None, 0, None, None)
        ref = a.double() @ w.double().t()
        a_hl = a_s[0].double() + a_s[1].double()
        w_hl = wp[0].double() + wp[1].double()
        ref_hl = a_hl.view(rows, K) @ w_hl.t()
        blas = (a @ w.t())
        scale = ref.abs().mean().item()
        for nm, got, rf in (("tc vs fp64(x)", out.view(rows, N).double(), ref), ("tc vs fp64(hi+lo)", out.view(rows, N).double(), ref_hl),
                            ("cublas fp32 vs fp64", blas.double(), ref)):
            e = got - rf
            print("%-9s %-20s mean|ref| %.3f  bias %.3e  rms %.3e  max %.3e  (relative to mean|ref|: bias %.2e rms %.2e)" %
                  (label, nm, scale, e.mean().item(), e.pow(2).mean().sqrt().item(), e.abs().max().item(),
                   e.mean().item() / scale, e.pow(2).mean().sqrt().item() / scale), flush=True)


def stage_gpu_noise():
    """The reference algorithm's own CPU-vs-GPU difference (oracle on cuda, TF32 off) for every end-to-end case."""
    import torch
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    import cases
    from oracle import unimatch_oracle as O
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    gold = torch.load(os.path.join(ROOT, "tests", "golden", "golden.pt"))["vectors"]
    for name in cases.E2E_CASES:
        cfg, sd, batch, call = cases.e2e_setup(name)
        sd = {k: v.cuda() for k, v in sd.items()}
        b = {k: v.cuda() for k, v in batch.items()}
        mk = {k: cfg["model"][k] for k in ("num_scales", "upsample_factor", "reg_refine")}
        out = O.forward(sd, b["img0"], b["img1"], intrinsics=b.get("intrinsics"), pose=b.get("pose"), **mk, **call)["flow_preds"][-1]
        mean, mx = cases.epe(out.cpu(), gold[name])
        print("%-28s oracle(cuda fp32) vs reference(cpu): mean %.3e max %.3e  (thread-noise %.1e, x%.1f)" %
              (name, mean, mx, cases.E2E_NOISE[name], mean / cases.E2E_NOISE[name]), flush=True)


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    if what == "all":
        for st in ("dump", "parity", "perf", "conv"):
            print("==== stage", st, flush=True)
            t0 = time.time()
            r = subprocess.run([sys.executable, os.path.abspath(__file__), st], timeout=600)
            print("==== stage %s rc=%d (%.1fs)" % (st, r.returncode, time.time() - t0), flush=True)
    else:
        {"dump": stage_dump, "parity": stage_parity, "perf": stage_perf, "conv": stage_conv, "backbone": stage_backbone, "accuracy": stage_accuracy, "gpu_noise": stage_gpu_noise}[what]()

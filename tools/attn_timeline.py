"""Pipeline timeline of the persistent attention kernel (CTA 0, its first two items).  Needs the diagnostic build of the
library (UM_ATTN_DEBUG_BUILD=1 in the environment: the stamp changes and the library is rebuilt with -DUM_ATTN_DEBUG=1).
Prints, per key tile, clock64 stamps relative to the first event: MMA warp (top of iteration, S issued, PV_A issued, PV_B issued)
and both softmax groups (wait start, S ready, S in registers, math done, P stored, arrive)."""
import ctypes
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("UM_ATTN_DBG", "32")
from unimatch_b200 import ops  # noqa: E402

OPS = torch.ops.unimatch_sm100
n = 16
h, w, K = (60, 104, 2) if "--s1" not in sys.argv else (120, 208, 8)
lp = ops.attention_planes_lp(h, w, K, K, 0, 0, 0)
lw = (h // K) * (w // K)
pl = []
for i in range(3):
    t = torch.zeros((2, n, K * K, lp, 128), device="cuda", dtype=torch.float16)
    t[:, :, :, :lw] = torch.randn((2, n, K * K, lw, 128), device="cuda").half()
    t[1] *= 1e-3
    pl.append(t)
out_s = torch.empty((2, n * h * w, 128), device="cuda", dtype=torch.float16)
dump = torch.zeros(24576, device="cuda")
f = lambda: OPS.window_attention_planes(pl[0], pl[1], pl[2], n, n // 2, h, w, K, K, 0, 0, 0, None, out_s)
f(); torch.cuda.synchronize()
ops.LIB.um_debug_set_dump(ctypes.c_void_p(dump.data_ptr()))
f(); torch.cuda.synchronize()
ops.LIB.um_debug_set_dump(None)
tl = dump.view(torch.int64).cpu()
T = (lw + 63) // 64
vals = tl[tl > 0]
t0 = int(vals.min())
rel = lambda v: (int(v) - t0) if v > 0 else -1
print("T =", T, "key tiles per item; cycles relative to the first stamp")
print("%4s | %28s | %44s | %44s" % ("tile", "MMA: top  S_iss PVA_iss PVB_iss", "softmax A: wait0 ready inreg math stored arr", "softmax B: ..."))
for jj in range(2 * T):
    m = [rel(tl[jj * 8 + e]) for e in range(4)]
    a = [rel(tl[1024 + jj * 8 + e]) for e in range(6)]
    b = [rel(tl[2048 + jj * 8 + e]) for e in range(6)]
    print("%4d | %6d %6d %6d %6d | %6d %6d %6d %6d %6d %6d | %6d %6d %6d %6d %6d %6d" % tuple([jj] + m + a + b))
for x in range(2):
    for k in range(2):
        e = [rel(tl[3072 + x * 16 + k * 4 + i]) for i in range(4)]
        print("epilogue group %d item %d: start %d  pv done %d  O in registers %d  stores issued %d" % (x, k, e[0], e[1], e[3], e[2]))

"""Launch list of a bench step -> shares per kernel (run here on the CSV the GPU box produced):
     ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches.csv \
         python bench.py --profile --steps 1
     python tools/launch_shares.py gpurun_out/launches.csv profiles/r02d_launch_shares.md
Only the LAST forward pass of the list is summarised (from the last image-stem launch on): the launches before it are the
weight preparation and the warm-up forward, which also fills the cached zero-padded buffers once.
Per-launch times under ncu are cold-cache and serialised: the SHARES are comparable with bench.py's CUDA-event shares, the
absolute values are not."""
import csv
import re
import sys

src, out = sys.argv[1], sys.argv[2]
rows = [r for r in csv.reader(l for l in open(src) if not l.startswith("==")) if r]
hdr = rows[0]
ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
body = rows[1:]
starts = [i for i, r in enumerate(body) if len(r) > ik and "stem7x7_kernel<3" in r[ik]]
skipped = starts[-1] if starts else 0
body = body[skipped:]
agg = {}
for r in body:
    if len(r) <= iv:
        continue
    name = r[ik]
    t = float(r[iv].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[iu], 1e-3)
    m = re.search(r"(\w+_kernel)(<[^>]*>)?", name)
    if "at::" in name or "elementwise" in name or "Cat" in name:
        m2 = re.search(r"(\w*elementwise_kernel|CatArrayBatchedCopy\w*|\w+_kernel)", name)
        f = re.search(r"(\w+Functor|\w+_kernel_cuda)", name)
        key = "ATen: %s%s" % (m2.group(1) if m2 else name[:40], " (%s)" % f.group(1) if f else "")
    else:
        key = (m.group(1) + (m.group(2) or "")) if m else name[:60]
    a = agg.setdefault(key, [0, 0.0])
    a[0] += 1
    a[1] += t
tot = sum(v[1] for v in agg.values())
own = sum(v[1] for k, v in agg.items() if not k.startswith("ATen"))
lib = [k for k in agg if re.search(r"cutlass|cublas|cudnn|sgemm|gemv", k, re.I)]
with open(out, "w") as f:
    f.write("# Launch shares: `ncu --metrics gpu__time_duration.sum --clock-control none` over `bench.py --profile --steps 1`\n\n")
    f.write("Last forward pass of the list (%d earlier launches = weight preparation + warm-up forward skipped).\n" % skipped)
    f.write("%d launches, %.1f ms summed (cold-cache, serialised per-launch times: shares are comparable with the CUDA-event shares of "
            "bench.py, absolutes are not).\nKernels of this repo: %.1f %% of the summed time; ATen glue: %.1f %%; cuBLAS / cuDNN / "
            "CUTLASS launches: %d.\n\n| kernel | launches | us | share |\n|---|---|---|---|\n"
            % (sum(v[0] for v in agg.values()), tot / 1e3, 100 * own / tot, 100 * (tot - own) / tot, len(lib)))
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write("| `%s` | %d | %.1f | %.2f %% |\n" % (k, v[0], v[1], 100 * v[1] / tot))
print("wrote", out)

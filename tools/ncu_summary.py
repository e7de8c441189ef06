"""Summarise an `ncu --set full` report (a .ncu-rep, or its `--page raw --csv` dump made on the GPU box when the report is too
large to bring back) into the JSON/markdown kept under profiles/ (run here, no GPU needed). Launches of ATen kernels (input
generation in the profiling scripts) are dropped; the labels name the remaining launches in order.
    python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r02_ncu_attention "label of launch 1" "label of launch 2" ..."""
import csv
import json
import subprocess
import sys

rep, out = sys.argv[1], sys.argv[2]
if rep.endswith(".csv"):
    raw = open(rep).read()
else:
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
rows = rows[:2] + [r for r in rows[2:] if "at::" not in r[idx["Kernel Name"]]]
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic"]
MULT = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1.0}
labels = sys.argv[3:]
items = []
for n, r in enumerate(rows[2:]):
    d = {"kernel": r[idx["Kernel Name"]].split("(")[0].split("::")[-1], "label": labels[n] if n < len(labels) else ""}
    for k in KEYS:
        if k in idx:
            v, u = r[idx[k]], units[idx[k]]
            try:
                f = float(v.replace(",", ""))
            except ValueError:
                continue
            if u in MULT and ("bytes" in k or "duration" in k):
                f *= MULT[u]
                u = "B" if "bytes" in k else "s"
            d[k] = f
            d[k + ".unit"] = u
    items.append(d)
json.dump(items, open(out + ".json", "w"), indent=1)
with open(out + ".md", "w") as f:
    f.write("# ncu --set full --clock-control none summaries (one launch each)\n\n")
    for d in items:
        f.write("## %s %s\n" % (d["kernel"], d["label"]))
        for k in KEYS:
            if k in d:
                f.write("- %s: %.6g %s\n" % (k, d[k], d[k + ".unit"]))
        f.write("\n")
print("wrote", out + ".json", out + ".md")

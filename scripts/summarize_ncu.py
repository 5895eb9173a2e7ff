"""Summarise ncu outputs into markdown for profiles/ (run in the build container; no GPU needed).

  python scripts/summarize_ncu.py launches gpurun_out/launches_r1.csv      > profiles/r1_launches.md
  python scripts/summarize_ncu.py report   gpurun_out/prof_r1.ncu-rep      > profiles/r1_kernels.md
"""
import csv
import io
import re
import subprocess
import sys
from collections import defaultdict


def short(name: str) -> str:
    name = re.sub(r"\(.*", "", name)
    name = re.sub(r"<.*", "", name)
    return name.split("::")[-1][:70]


def launches(path):
    rows = []
    with open(path, newline="") as f:
        lines = [l for l in f if not l.startswith("==")]
    rd = csv.DictReader(io.StringIO("".join(lines)))
    for r in rd:
        if r.get("Metric Name") == "gpu__time_duration.sum":
            v = float(r["Metric Value"].replace(",", ""))
            unit = r.get("Metric Unit", "ns")
            scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1e-3)
            rows.append((short(r["Kernel Name"]), v * scale))
    tot = sum(t for _, t in rows)
    agg = defaultdict(lambda: [0, 0.0])
    for k, t in rows:
        agg[k][0] += 1
        agg[k][1] += t
    ours = ("attn_prefill", "attn_decode", "decode_stack", "gemv_kernel", "resample_", "kv_", "rmsnorm", "swiglu", "argmax", "splice_rows", "gather_rows", "dedup", "voxel",
            "compact_")
    print(f"{len(rows)} kernel launches, {tot / 1e3:.2f} ms of kernel time (ncu serialises launches and flushes caches: compare SHARES)\n")
    print("| kernel | launches | total us | avg us | share | ours |")
    print("|---|---|---|---|---|---|")
    mine = 0.0
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
        is_ours = any(o in k for o in ours)
        mine += t if is_ours else 0
        print(f"| {k} | {n} | {t:.1f} | {t / n:.2f} | {100 * t / tot:.2f}% | {'yes' if is_ours else ''} |")
    print(f"\nlibsvln kernels: {100 * sum(t for k, (n, t) in agg.items() if any(o in k for o in ours)) / tot:.2f}% of kernel time")


WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "smsp__cycles_active.avg", "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_uniform.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]


def report(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rd = list(csv.reader(io.StringIO(out)))
    hdr, units = rd[0], rd[1]
    col = {h: i for i, h in enumerate(hdr)}
    avail = [w for w in WANT if w in col]
    tens = [h for h in hdr if "tensor" in h and "pct" in h][:6]
    cols = avail + [t for t in tens if t not in avail]
    print("| kernel | " + " | ".join(c.replace("__", " ").replace(".avg.pct_of_peak_sustained_", " %") for c in cols) + " |")
    print("|---|" + "---|" * len(cols))
    for r in rd[2:]:
        if len(r) < len(hdr):
            continue
        print(f"| {short(r[col['Kernel Name']])} | " + " | ".join(f"{r[col[c]]} {units[col[c]]}".strip() for c in cols) + " |")


if __name__ == "__main__":
    {"launches": launches, "report": report}[sys.argv[1]](sys.argv[2])

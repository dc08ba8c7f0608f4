"""Summaries of ncu captures for profiles/: (1) launch list CSV (`ncu --metrics gpu__time_duration.sum --csv`) ->
per-kernel totals and shares; (2) `--set full` report (.ncu-rep) -> the handful of metrics DESIGN.md quotes.
Usage: python scripts/ncu_summary.py launches <launches.csv> <out.csv>
       python scripts/ncu_summary.py full <report.ncu-rep> <out.json>"""
import collections
import csv
import json
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_tag_requests.avg.pct_of_peak_sustained_elapsed",
        "sm__cycles_elapsed.max", "launch__block_size", "launch__grid_size",
        "launch__shared_mem_per_block_dynamic"]


def launches(src, dst):
    rows = list(csv.reader(open(src)))
    start = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    hdr = rows[start]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[start + 1:]:
        if len(r) <= vi or not r[vi]:
            continue
        v = float(r[vi].replace(",", ""))
        if v != v:  # ncu reports "nan" for some launches in the single-metric pass: counted, not timed
            agg.setdefault(r[ki].split("(")[0] + " [duration not reported by ncu]", []).append(0.0)
            continue
        v = v / 1e3 if r[ui] in ("ns", "nsecond") else v  # -> us
        name = r[ki].split("(")[0]
        agg.setdefault(name, []).append(v)
    total = sum(sum(v) for v in agg.values())
    with open(dst, "w") as f:
        f.write("kernel,launches,total_ms,avg_us,share\n")
        for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
            f.write(f"\"{k}\",{len(v)},{sum(v) / 1e3:.3f},{sum(v) / len(v):.1f},{sum(v) / total:.4f}\n")
    print(open(dst).read())


def full(src, dst):
    raw = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    out, seen = [], set()
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        d = {"Kernel Name": name}
        for k in KEYS:
            if k in hdr:
                d[k] = [r[hdr.index(k)], units[hdr.index(k)]]
        key = (name, d.get("launch__grid_size", [""])[0], d.get("launch__shared_mem_per_block_dynamic", [""])[0])
        if key in seen:
            continue
        seen.add(key)
        out.append(d)
    json.dump(out, open(dst, "w"), indent=1)
    for d in out:
        print(d["Kernel Name"][:60], d.get("gpu__time_duration.sum"), d.get("dram__bytes_read.sum"), d.get("dram__bytes_write.sum"))


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3])

#!/usr/bin/env python
"""Summarise an ncu `--set full` capture (an .ncu-rep, or the CSV of its raw page) into one line per captured launch.

    python tools/ncu_summary.py file.ncu-rep|file_raw.csv [--md] [--merge]

--md prints a markdown table; --merge averages launches of the same kernel at the same grid size.
"""
import collections
import csv
import io
import subprocess
import sys

KEYS = [("gpu__time_duration.sum", "us"), ("dram__bytes_read.sum", "rdMB"), ("dram__bytes_write.sum", "wrMB"),
        ("FBSP.TriageCompute.dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
        ("lts__t_sector_hit_rate.pct", "L2hit%"), ("l1tex__t_sector_hit_rate.pct", "L1hit%"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor%"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts%"),
        ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex%"),
        ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "lsu%"),
        ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "fma%"),
        ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "alu%"),
        ("lts__t_sectors_op_red.sum", "red_sectors"),
        ("l1tex__t_requests_pipe_lsu_mem_global_op_red.sum", "red_reqs"),
        ("launch__registers_per_thread", "regs"), ("launch__grid_size", "grid")]
STALL = "smsp__average_warps_issue_stalled_"
MD_COLS = ["us", "rdMB", "wrMB", "dram%", "L2hit%", "tensor%", "l1tex%", "lts%", "issue%", "occ%", "regs", "grid"]


def load(path):
    if path.endswith(".ncu-rep"):
        text = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    else:
        text = open(path).read()
    rows = list(csv.reader(io.StringIO(text)))
    return rows[0], rows[1], rows[2:]


def launches(path):
    hdr, units, rows = load(path)
    col = {k: i for i, k in enumerate(hdr)}
    stall_cols = [(k[len(STALL):-len("_per_issue_active.ratio")], i) for k, i in col.items()
                  if k.startswith(STALL) and k.endswith("_per_issue_active.ratio") and "not_selected" not in k and
                  "_selected" not in k]
    out = []
    for r in rows:
        name = r[col["Kernel Name"]].replace("void ", "").replace("<unnamed>::", "").split("(")[0][:44]
        m = collections.OrderedDict()
        for k, label in KEYS:
            if k not in col or label in m:
                continue
            v, u = r[col[k]], units[col[k]]
            try:
                f = float(v.replace(",", ""))
            except ValueError:
                continue
            if label == "us":
                f = f / 1e3 if u in ("ns", "nsecond") else (f * 1e3 if u in ("ms", "msecond") else f)
            if label.endswith("MB"):
                f = {"byte": f / 1e6, "Kbyte": f / 1e3, "Mbyte": f, "Gbyte": f * 1e3}.get(u, f)
            m[label] = f
        st = []
        for s, i in stall_cols:
            try:
                st.append((float(r[i].replace(",", "")), s))
            except ValueError:
                pass
        st.sort(reverse=True)
        tot = sum(v for v, _ in st) or 1.0
        m["stall"] = ", ".join(f"{s} {100 * v / tot:.0f}%" for v, s in st[:2])
        out.append((name, m))
    return out


def merge(ls):
    groups = collections.OrderedDict()
    for name, m in ls:
        groups.setdefault((name, m.get("grid")), []).append(m)
    out = []
    for (name, _), ms in groups.items():
        avg = collections.OrderedDict()
        for k in ms[0]:
            avg[k] = ms[0][k] if isinstance(ms[0][k], str) else sum(x.get(k, 0.0) for x in ms) / len(ms)
        avg["n"] = len(ms)
        out.append((name, avg))
    return out


def main():
    ls = launches(sys.argv[1])
    if "--merge" in sys.argv:
        ls = merge(ls)
    if "--md" in sys.argv:
        cols = (["n"] if "--merge" in sys.argv else []) + MD_COLS
        print("| kernel | " + " | ".join(cols) + " | top stalls (share of stalled warp-cycles) |")
        print("|---|" + "---|" * (len(cols) + 1))
        for name, m in ls:
            print(f"| `{name}` | " + " | ".join(f"{m[c]:.4g}" if c in m else "–" for c in cols) + f" | {m['stall']} |")
    else:
        for name, m in ls:
            print(name, " ".join(f"{k}={v:.4g}" if not isinstance(v, str) else f"{k}=[{v}]" for k, v in m.items()))


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Summarise an .ncu-rep (raw page) into one line per captured launch: python tools/ncu_summary.py file.ncu-rep"""
import csv
import io
import subprocess
import sys

KEYS = [("gpu__time_duration.sum", "us"), ("dram__bytes_read.sum", "rdMB"), ("dram__bytes_write.sum", "wrMB"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
        ("lts__t_sector_hit_rate.pct", "L2hit%"), ("l1tex__t_sector_hit_rate.pct", "L1hit%"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor%"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"),
        ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "lsu%"),
        ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "xu%"),
        ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "fma%"),
        ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "alu%"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
        ("lts__t_sectors_op_red.sum", "red_sectors"), ("lts__t_sectors_op_atom.sum", "atom_sectors"),
        ("l1tex__t_requests_pipe_lsu_mem_global_op_red.sum", "red_reqs"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts%"),
        ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex%"),
        ("launch__registers_per_thread", "regs"), ("launch__grid_size", "grid")]


def main():
    out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        name = name.replace("void ", "").replace("<unnamed>::", "").split("(")[0][:48]
        parts = []
        for k, label in KEYS:
            if k in hdr:
                v, u = r[hdr.index(k)], units[hdr.index(k)]
                try:
                    f = float(v.replace(",", ""))
                    if label == "us":
                        f = f / 1e3 if u in ("ns", "nsecond") else (f * 1e3 if u in ("ms", "msecond") else f)
                    if label.endswith("MB"):
                        f = {"byte": f / 1e6, "Kbyte": f / 1e3, "Mbyte": f, "Gbyte": f * 1e3}.get(u, f)
                    parts.append(f"{label}={f:.4g}")
                except ValueError:
                    parts.append(f"{label}={v}")
        print(name, " ".join(parts))


if __name__ == "__main__":
    main()

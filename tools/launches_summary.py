#!/usr/bin/env python
"""Turn the ncu launch list of one bench step (`ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,
dram__bytes_write.sum --csv`) into a per-kernel table (markdown) and the DRAM traffic of the dense-layer kernels
(profiles/traffic.json, read by bench.py for roofline.traffic).

    python tools/launches_summary.py gpurun_out/launches.csv profiles/r01_launches.md [--precision bf16 --grad full]
"""
import collections
import csv
import json
import os
import re
import sys

DENSE = ("tc_gemm_nt", "tc_gemm_tn", "tc_heads", "tn_reduce", "tn_colsum_reduce", "rowdot", "gemm_nt_kernel", "gemm_tn_kernel",
         "sdf_trunk")  # kernels behind the dense-layer entry points counted in bench.py's roofline


def main():
    src, dst = sys.argv[1], sys.argv[2]
    prec = sys.argv[sys.argv.index("--precision") + 1] if "--precision" in sys.argv else "bf16"
    grad = sys.argv[sys.argv.index("--grad") + 1] if "--grad" in sys.argv else "full"
    rows = list(csv.reader(open(src)))
    hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
    hdr = rows[hi]
    kn, mn, mv, mu, idc = (hdr.index(x) for x in ("Kernel Name", "Metric Name", "Metric Value", "Metric Unit", "ID"))
    d = collections.OrderedDict()
    for r in rows[hi + 1:]:
        if len(r) <= mv:
            continue
        val = float(r[mv].replace(",", ""))
        unit = r[mu]
        if r[mn].startswith("gpu__time"):
            val *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3}.get(unit, 1.0)
        else:
            val *= {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)
        d.setdefault((int(r[idc]), r[kn]), {})[r[mn]] = val
    seq = []
    for (i, n), m in d.items():
        name = re.sub(r"void |<unnamed>::|\(.*", "", n)
        seq.append((name, m.get("gpu__time_duration.sum", 0.0),
                    m.get("dram__bytes_read.sum", 0.0) + m.get("dram__bytes_write.sum", 0.0)))
    starts = [k for k, s in enumerate(seq) if "rays_from_pose" in s[0]]
    # one step = the launches between two ray-generation kernels; prefer a slice of the plain fwd+bwd loop (bench.py
    # also runs a loop with the optimizer step inside)
    spans = list(zip(starts, starts[1:] + [len(seq)]))
    plain = [sp for sp in spans[:-1] if not any("adamw" in n for n, _, _ in seq[sp[0]:sp[1]])] or spans[:1]
    a, b = plain[0]
    opt = [x for x in seq[a:b] if "adamw" in x[0]]
    step = [x for x in seq[a:b] if "adamw" not in x[0]]   # the bench value times fwd+bwd; the optimizer is reported apart
    agg = collections.defaultdict(lambda: [0, 0.0, 0.0])
    for n, t, by in step:
        agg[n][0] += 1
        agg[n][1] += t
        agg[n][2] += by
    total = sum(t for _, t, _ in step)
    lines = [f"One training step = {len(step)} launches, {total:.0f} us summed device time (ncu: serialised, cold cache;"
             " compare shares, not absolutes).", "",
             "| kernel | launches | us | share | DRAM MB | GB/s |", "|---|---|---|---|---|---|"]
    for n, (c, t, by) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        lines.append(f"| `{n[:70]}` | {c} | {t:.1f} | {100 * t / total:.1f}% | {by / 1e6:.1f} | {by / t / 1e3 if t else 0:.0f} |")
    if opt:
        lines += ["", "The captured slice came from bench.py's with-optimizer loop; its optimizer launch is not part of the "
                  "step above: " + ", ".join(f"`{n}` {t:.1f} us, {by / 1e6:.1f} MB DRAM ({by / t / 1e3:.0f} GB/s)"
                                             for n, t, by in opt) + "."]
    open(dst, "w").write("\n".join(lines) + "\n")
    dense_bytes = sum(by for n, t, by in step if any(k in n for k in DENSE))
    dense_us = sum(t for n, t, by in step if any(k in n for k in DENSE))
    out = {"precision": prec, "grad": grad, "dense_layers_dram_bytes_per_step": dense_bytes,
           "dense_layers_us_per_step_ncu": dense_us, "step_us_ncu": total, "source": os.path.basename(src)}
    json.dump(out, open(os.path.join(os.path.dirname(dst), "traffic.json"), "w"), indent=1)
    print("\n".join(lines[:12]))
    print(out)


if __name__ == "__main__":
    main()

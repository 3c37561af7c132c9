#!/bin/bash
# tools/profile_step.sh -- the ncu evidence of one round (run under gpurun, ONE GPU): launch list of a training step,
# `--set full` captures of the step's top kernels and of the fused head forward, L2-window A/B.  Every ncu command is
# preceded by the same command without ncu (B200_PROFILING.md).
R=${1:-r02}
L2AB=${2:-yes}   # second argument "no": skip the L2-window A/B (three more bench runs)
O=gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-graph --no-extras --no-cpu-baseline"
$CMD > $O/${R}_plain.log 2>&1 &&
timeout 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 700 -c 420 \
    --csv --log-file $O/${R}_launches.csv $CMD > $O/${R}_ncu_launches.log 2>&1
echo "launch list rc=$?"
K='regex:tc_gemm_nt_persist|tc_gemm_tn_kernel|tc_sdf_trunk_fused|encode_rays_tcl|encode_rays_bwd_tcl|tc_heads_kernel|sdf_trunk_bwd_kernel|tc_gemm_nt_kernel|sample_merge_fine'
$CMD > $O/${R}_plain2.log 2>&1 &&
timeout 600 ncu --set full --clock-control none -k "$K" -s 190 -c 40 -o $O/${R}_step_full $CMD > $O/${R}_ncu_full.log 2>&1
echo "full rc=$?"
# gpurun copies at most 64 MiB back: keep the raw page as CSV, drop the report
ncu -i $O/${R}_step_full.ncu-rep --page raw --csv > $O/${R}_step_full_raw.csv 2>/dev/null && rm -f $O/${R}_step_full.ncu-rep
H="python tools/bench_heads.py --only fwd_nostore --iters 3"
$H > /dev/null 2>&1 &&
timeout 250 ncu --set full --clock-control none -k regex:tc_heads -s 2 -c 1 -o $O/${R}_heads_fwd_nostore $H > $O/${R}_ncu_heads.log 2>&1
echo "heads rc=$?"
ncu -i $O/${R}_heads_fwd_nostore.ncu-rep --page raw --csv > $O/${R}_heads_fwd_nostore_raw.csv 2>/dev/null && rm -f $O/${R}_heads_fwd_nostore.ncu-rep
B="python bench.py --steps 20 --warmup 3 --no-extras --no-cpu-baseline"
[ "$L2AB" = "no" ] && { du -sh $O | tail -1; exit 0; }
for hr in 0 0.6 1.0; do
  MLI_L2_PERSIST=$hr timeout 200 $B > $O/${R}_l2_$hr.json 2> $O/${R}_l2_$hr.err || tail -3 $O/${R}_l2_$hr.err
  python - <<PY
import json
d = json.loads([l for l in open("$O/${R}_l2_$hr.json") if l.startswith("{")][-1])
enc = [k for k in d["roofline"]["kernels"] if k["entry"] in ("mli_encode_rays_tcl", "mli_encode_rays_bwd_tcl")]
print("MLI_L2_PERSIST=$hr", "ms/step", round(d["ms_per_step"], 4), {k["entry"]: k["us_per_step"] for k in enc})
PY
done
python -c "
from mli_nerf_b200 import _lib
print('L2 bytes / max persisting / max window:', _lib.l2_info())"
du -sh $O | tail -1

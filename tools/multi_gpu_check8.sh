#!/bin/bash
# tools/multi_gpu_check8.sh N -- trimmed variant of multi_gpu_check.sh for large N (box time is charged N x): the
# reduce-scatter exchange over NCCL (device-path check + bench with the extra sections) and the all-reduce bench for A/B.
N=${1:-8}
OUT=gpurun_out/multi_n${N}
mkdir -p gpurun_out
: > $OUT.log
run() { echo "== $*" | tee -a $OUT.log; timeout 200 "$@" >> $OUT.log 2>&1; echo "rc=$?" | tee -a $OUT.log; }
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node=$N --master-addr 127.0.0.1"
MLI_TABLE_EXCHANGE=reduce_scatter run $TR --master-port 29734 tests/multi/train_step_check.py
MLI_TABLE_EXCHANGE=allreduce run $TR --master-port 29736 tests/multi/train_step_check.py
run $TR --master-port 29735 tests/multi/inference_shard_check.py
grep -E "_OK|rc=|Error" $OUT.log | head -20
for table in reduce_scatter allreduce; do
  echo "== bench --gpus $N exchange=$table" | tee -a $OUT.log
  EXTRA=""; [ "$table" = "allreduce" ] && EXTRA="--no-extras"
  MLI_TABLE_EXCHANGE=$table timeout 300 $TR --master-port 29740 bench.py --gpus $N --steps 30 --warmup 5 --no-cpu-baseline $EXTRA \
      > ${OUT}_bench_$table.json 2>> $OUT.log
  echo "rc=$?" | tee -a $OUT.log
  python - <<PY
import json
try:
    d = json.loads([l for l in open("${OUT}_bench_$table.json") if l.startswith("{")][-1])
    print("$table", "N=", d["n_gpus"], "rays/s", round(d["value"]), "ms", round(d["ms_per_step"], 3), "host ms", round(d["host_enqueue_ms_per_step"], 2),
          "with_opt ms", round(d["with_optimizer"]["ms_per_step"], 3), d["config"]["exchange"], "render", d.get("render", {}).get("rays_per_sec"),
          {k: round(v["value"]) for k, v in d.get("workloads", {}).items()})
except Exception as e:
    print("$table", "no bench line:", e)
PY
done
echo "== bench --gpus $N exchange=reduce_scatter MLI_COMM_SMS=16" | tee -a $OUT.log
MLI_COMM_SMS=16 timeout 200 $TR --master-port 29741 bench.py --gpus $N --steps 30 --warmup 5 --no-cpu-baseline --no-extras \
    > ${OUT}_bench_rs_sms16.json 2>> $OUT.log
python - <<PY
import json
try:
    d = json.loads([l for l in open("${OUT}_bench_rs_sms16.json") if l.startswith("{")][-1])
    print("rs comm_sms=16", "rays/s", round(d["value"]), "ms", round(d["ms_per_step"], 3), "host ms", round(d["host_enqueue_ms_per_step"], 2))
except Exception as e:
    print("no bench line:", e)
PY
tail -4 $OUT.log

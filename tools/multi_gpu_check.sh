#!/bin/bash
# tools/multi_gpu_check.sh N -- multi-GPU device-path checks + exchange A/B on an N-GPU box (gpurun --gpus N):
#   tests/multi/*.py under torchrun (both exchanges x both transports), then bench.py --gpus N for each exchange.
N=${1:-2}
OUT=gpurun_out/multi_n${N}
mkdir -p gpurun_out
: > $OUT.log
run() { echo "== $*" | tee -a $OUT.log; timeout 240 "$@" >> $OUT.log 2>&1; echo "rc=$?" | tee -a $OUT.log; }
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node=$N --master-addr 127.0.0.1"
run $TR --master-port 29733 tests/multi/peer_allreduce_check.py
for table in allreduce reduce_scatter; do
  for mode in peer nccl; do
    MLI_TABLE_EXCHANGE=$table MLI_TABLE_ALLREDUCE=$mode run $TR --master-port 29734 tests/multi/train_step_check.py
  done
done
run $TR --master-port 29735 tests/multi/inference_shard_check.py
grep -E "_OK|rc=" $OUT.log
for table in reduce_scatter allreduce; do
  echo "== bench --gpus $N exchange=$table" | tee -a $OUT.log
  EXTRA=""; [ "$table" = "allreduce" ] && EXTRA="--no-extras"
  MLI_TABLE_EXCHANGE=$table timeout 400 $TR --master-port 29740 bench.py --gpus $N --steps 30 --warmup 5 --no-cpu-baseline $EXTRA \
      > ${OUT}_bench_$table.json 2>> $OUT.log
  echo "rc=$?" | tee -a $OUT.log
  python - <<PY
import json
try:
    d = json.load(open("${OUT}_bench_$table.json"))
    print("$table", "N=", d["n_gpus"], "rays/s", round(d["value"]), "ms", round(d["ms_per_step"], 3), "host ms", round(d["host_enqueue_ms_per_step"], 2),
          "with_opt ms", round(d["with_optimizer"]["ms_per_step"], 3), d["config"]["exchange"], "render", d.get("render", {}).get("rays_per_sec"))
except Exception as e:
    print("$table", "no bench line:", e)
PY
done
tail -5 $OUT.log

#!/bin/bash
# tools/multi_gpu_quick.sh N TAG -- the short form of multi_gpu_check.sh: device-path checks of the default exchange on both
# transports + the reference all-reduce on NCCL + sharded inference, then ONE bench line (reduce-scatter, no extras).
N=${1:-2}
TAG=${2:-quick}
OUT=gpurun_out/multi_n${N}_${TAG}
mkdir -p gpurun_out
: > $OUT.log
run() { echo "== $*" | tee -a $OUT.log; timeout 240 "$@" >> $OUT.log 2>&1; echo "rc=$?" | tee -a $OUT.log; }
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node=$N --master-addr 127.0.0.1"
MLI_TABLE_EXCHANGE=reduce_scatter MLI_TABLE_ALLREDUCE=nccl run $TR --master-port 29734 tests/multi/train_step_check.py
[ "$N" = "2" ] && MLI_TABLE_EXCHANGE=reduce_scatter MLI_TABLE_ALLREDUCE=peer run $TR --master-port 29734 tests/multi/train_step_check.py
MLI_TABLE_EXCHANGE=allreduce MLI_TABLE_ALLREDUCE=nccl run $TR --master-port 29734 tests/multi/train_step_check.py
run $TR --master-port 29735 tests/multi/inference_shard_check.py
grep -E "_OK|rc=" $OUT.log
MLI_TABLE_EXCHANGE=reduce_scatter timeout 400 $TR --master-port 29740 bench.py --gpus $N --steps 30 --warmup 5 --no-cpu-baseline --no-extras \
    > ${OUT}_bench.json 2>> $OUT.log
echo "bench rc=$?" | tee -a $OUT.log
python - <<PY
import json
try:
    d = json.loads([l for l in open("${OUT}_bench.json") if l.startswith("{")][-1])
    print("N=", d["n_gpus"], "rays/s", round(d["value"]), "ms", round(d["ms_per_step"], 3), "host ms", round(d["host_enqueue_ms_per_step"], 2),
          "with_opt ms", round(d["with_optimizer"]["ms_per_step"], 3), d["config"]["exchange"])
except Exception as e:
    print("no bench line:", e)
PY
tail -3 $OUT.log

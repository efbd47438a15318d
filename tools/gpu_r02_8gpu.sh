#!/bin/bash
# 8-GPU call (charged 8x: keep it short): NCCL world-3 tests, the contract bench at 8 and 4 GPUs, BASELINE configs[4]
# (the --optimize sweep on 800M f64 keys over 8 replicas).
tag=${1:-r02j}
out=gpurun_out
mkdir -p $out
export PYTHONUNBUFFERED=1
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm --format=csv > $out/${tag}_smi.txt 2>&1
timeout 300 python -m pytest tests/test_gpu_sharded.py -x -q -m gpu -k "3-uniform-linear,linear or 3-dups-cubic or 3-lognormal-lognormal" > $out/${tag}_pytest_sharded.log 2>&1
echo "pytest sharded exit $?" >> $out/${tag}_pytest_sharded.log; tail -3 $out/${tag}_pytest_sharded.log
for n in 8 4; do
  NCCL_DEBUG=WARN timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) \
     bench.py --gpus $n --steps 20 --warmup 3 > $out/${tag}_bench_n$n.json 2> $out/${tag}_bench_n$n.err
  echo "bench N=$n exit $?"
  python - <<PY
import json
try:
    d=json.load(open("$out/${tag}_bench_n$n.json"))
    print("N=$n ms", d["ms_per_step"], "value", d["value"], d.get("parity"), d["roofline"]["phases_ms"], "e2e", d["e2e"]["ms_per_step"])
    for k,v in d["extra_configs"].items(): print("  ", k, v.get("ms_per_step"), v.get("phases_ms"), v.get("device_ms"))
except Exception as e:
    print("no bench line", e)
PY
  tail -2 $out/${tag}_bench_n$n.err
done
timeout 600 python tools/optimize_bench.py --keys 800e6 --gpus 8 --sample 2e6 --skip-single > $out/${tag}_optimize_8gpu.json 2> $out/${tag}_optimize_8gpu.err
echo "optimize exit $?"; cut -c1-700 $out/${tag}_optimize_8gpu.json; tail -3 $out/${tag}_optimize_8gpu.err
RMI_OPTIMIZER_NO_BATCH=1 timeout 600 python tools/optimize_bench.py --keys 800e6 --gpus 8 --sample 0 --skip-single > $out/${tag}_optimize_8gpu_nobatch.json 2> $out/${tag}_optimize_8gpu_nobatch.err
echo "optimize (unbatched) exit $?"; cut -c1-300 $out/${tag}_optimize_8gpu_nobatch.json

#!/bin/bash
# 8-GPU call: timeline of the one-call range-partitioned build at 8 ranks, then the contract bench at N = 8 and 4.
#   gpurun --gpus 8 -- 'bash tools/gpu_r02_scale.sh <tag>'
tag=${1:-r02u}
out=gpurun_out; mkdir -p $out
export PYTHONUNBUFFERED=1
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm --format=csv > $out/${tag}_smi.txt 2>&1
RMI_DEV_SHARD_TRACE=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29611 \
   bench.py --gpus 8 --steps 6 --warmup 3 --no-extras --no-cpu-baseline > $out/${tag}_trace_bench_n8.json 2> $out/${tag}_trace_n8.err
grep "shard trace" $out/${tag}_trace_n8.err | tail -16 | cut -c1-600
for n in 8 4; do
  NCCL_DEBUG=WARN timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) \
     bench.py --gpus $n --steps 20 --warmup 3 > $out/${tag}_bench_n$n.json 2> $out/${tag}_bench_n$n.err
  echo "bench N=$n exit $?"
  python - <<PY
import json
d=json.loads(open("$out/${tag}_bench_n$n.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("value","ms_per_step","n_gpus")}, d["e2e"], d["parity"], d["roofline"]["phases_ms"])
print(json.dumps(d["extra_configs"])[:1200])
PY
  tail -2 $out/${tag}_bench_n$n.err
done

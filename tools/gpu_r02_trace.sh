#!/bin/bash
# Multi-GPU developer call: timeline of the one-call range-partitioned build (RMI_DEV_SHARD_TRACE) at N ranks, then a pytest subset.
#   gpurun --gpus N -- 'bash tools/gpu_r02_trace.sh <tag> <N> [pytest -k expression]'
tag=${1:-r02s}; N=${2:-2}; K=$3
out=gpurun_out; mkdir -p $out
export PYTHONUNBUFFERED=1
RMI_DEV_SHARD_TRACE=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 \
   bench.py --gpus $N --steps 6 --warmup 3 --no-extras --no-cpu-baseline > $out/${tag}_trace_bench_n$N.json 2> $out/${tag}_trace_n$N.err
grep "shard trace" $out/${tag}_trace_n$N.err | tail -$((2 * N)) | cut -c1-700
if [ -n "$K" ]; then
  timeout 900 python -m pytest tests/test_gpu_sharded.py -x -q -m gpu -k "$K" > $out/${tag}_pytest_sharded.log 2>&1
  echo "pytest sharded exit $?" >> $out/${tag}_pytest_sharded.log; tail -4 $out/${tag}_pytest_sharded.log
fi

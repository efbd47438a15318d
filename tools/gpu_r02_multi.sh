#!/bin/bash
# Multi-GPU call: NCCL tests of the range-partitioned build, the contract bench at N GPUs, the optimizer sweep on replicas.
#   gpurun --gpus N -- 'bash tools/gpu_r02_multi.sh <tag> <N> [optimize_keys]'
tag=${1:-r02m}
N=${2:-2}
okeys=${3:-200e6}
out=gpurun_out
mkdir -p $out
export PYTHONUNBUFFERED=1
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm --format=csv > $out/${tag}_smi.txt 2>&1
nvidia-smi topo -m >> $out/${tag}_smi.txt 2>&1
[ "$4" = shardtime ] && timeout 300 python tools/dev_shard_time.py > $out/${tag}_shard_time.json 2> $out/${tag}_shard_time.err; cat $out/${tag}_shard_time.json
timeout 900 python -m pytest tests/test_gpu_sharded.py -x -q -m gpu ${PYTEST_K:+-k "$PYTEST_K"} > $out/${tag}_pytest_sharded.log 2>&1
echo "pytest sharded exit $?" >> $out/${tag}_pytest_sharded.log
tail -4 $out/${tag}_pytest_sharded.log
for n in $(seq 2 $N | awk -v N=$N '{ if ($1==2 || $1==4 || $1==8) print $1 }'); do
  NCCL_DEBUG=WARN timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) \
     bench.py --gpus $n --steps 20 --warmup 3 > $out/${tag}_bench_n$n.json 2> $out/${tag}_bench_n$n.err
  echo "bench N=$n exit $?"
  cut -c1-2500 $out/${tag}_bench_n$n.json
  tail -3 $out/${tag}_bench_n$n.err
done
if [ "$okeys" != "0" ]; then
timeout 900 python tools/optimize_bench.py --keys $okeys --gpus $N --sample 2e6 > $out/${tag}_optimize_${N}gpu.json 2> $out/${tag}_optimize_${N}gpu.err
echo "optimize exit $?"; cut -c1-1500 $out/${tag}_optimize_${N}gpu.json; tail -3 $out/${tag}_optimize_${N}gpu.err
RMI_OPTIMIZER_NO_BATCH=1 timeout 900 python tools/optimize_bench.py --keys $okeys --gpus $N --sample 0 > $out/${tag}_optimize_${N}gpu_nobatch.json 2> $out/${tag}_optimize_${N}gpu_nobatch.err
echo "optimize (unbatched) exit $?"; cut -c1-600 $out/${tag}_optimize_${N}gpu_nobatch.json
fi

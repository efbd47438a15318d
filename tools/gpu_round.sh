#!/bin/bash
# One gpurun call: GPU parity tests, smoke, the contract bench, the ncu launch list and one full
# capture of the dominant kernel.  Everything lands in gpurun_out/<tag>_*.
# usage: tools/gpu_round.sh <tag> [skip-ncu | launches-only]   (launches-only: no dev_bench, no full capture)
tag=${1:-run}
out=gpurun_out
mkdir -p $out
export PYTHONUNBUFFERED=1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $out/${tag}_smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > $out/${tag}_pytest_gpu.log 2>&1
echo "pytest exit $?" >> $out/${tag}_pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > $out/${tag}_smoke.log 2>&1
echo "smoke exit $?" >> $out/${tag}_smoke.log
timeout 600 python bench.py > $out/${tag}_bench_n1.json 2> $out/${tag}_bench_n1.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $out/${tag}_bench_ref.json 2> $out/${tag}_bench_ref.err
if [ "$2" != "launches-only" ]; then
  timeout 300 python tools/dev_bench.py --iters=3 > $out/${tag}_dev_bench.jsonl 2>&1
fi
if [ "$2" != "skip-ncu" ]; then
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
     --log-file $out/${tag}_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $out/${tag}_ncu_bench.log 2>&1
  if [ "$2" != "launches-only" ]; then
    timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_leaf -s 10 -c 5 \
       -o $out/${tag}_k_leaf_full -f python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $out/${tag}_ncu_full.log 2>&1
  fi
fi
tail -3 $out/${tag}_pytest_gpu.log
tail -2 $out/${tag}_smoke.log
cat $out/${tag}_bench_n1.json

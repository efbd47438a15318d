#!/bin/bash
# One-GPU call on the final sources: the whole GPU test suite, smoke, the contract bench (both arms), the ncu launch list and
# one full capture of the dominant kernel.
tag=${1:-r02h}
out=gpurun_out
mkdir -p $out
export PYTHONUNBUFFERED=1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $out/${tag}_smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > $out/${tag}_pytest_gpu.log 2>&1
echo "pytest exit $?" >> $out/${tag}_pytest_gpu.log
tail -3 $out/${tag}_pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > $out/${tag}_smoke.log 2>&1
echo "smoke exit $?" >> $out/${tag}_smoke.log
tail -2 $out/${tag}_smoke.log
timeout 900 python bench.py > $out/${tag}_bench_n1.json 2> $out/${tag}_bench_n1.err
echo "bench exit $?"; cut -c1-600 $out/${tag}_bench_n1.json
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > $out/${tag}_bench_ref.json 2> $out/${tag}_bench_ref.err
echo "reference arm exit $?"; cut -c1-500 $out/${tag}_bench_ref.json
timeout 300 python tools/dev_bench.py --iters=5 > $out/${tag}_dev_bench.jsonl 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
   --log-file $out/${tag}_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras > $out/${tag}_ncu_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_leaf -s 10 -c 6 \
   -o $out/${tag}_k_leaf_full -f python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras > $out/${tag}_ncu_full.log 2>&1
timeout 600 python tools/optimize_bench.py --keys 200e6 --gpus 1 --sample 0 > $out/${tag}_optimize_1gpu.json 2> $out/${tag}_optimize_1gpu.err
echo "optimize exit $?"; cut -c1-400 $out/${tag}_optimize_1gpu.json

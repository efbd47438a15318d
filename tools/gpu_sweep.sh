#!/bin/bash
# Knob sweep of the leaf kernel on the bench workload (200M uint64, linear,linear 2^20):
# L2 eviction hints of the two passes (RMI_DEV_L2_HINT=<fit><fwd>, 0 normal 1 evict_first 2 evict_last)
# and the number of launch slices whose results are copied to the host while later slices compute.
tag=${1:-sweep}
out=gpurun_out
mkdir -p $out
export PYTHONUNBUFFERED=1
f=$out/${tag}_sweep.jsonl
: > $f
for hint in 00 21 01 20 22; do
  RMI_DEV_L2_HINT=$hint timeout 200 python tools/dev_bench.py --one --iters=8 2>&1 | tail -1 >> $f
done
for sl in 1 4 16; do
  RMI_DEV_LEAF_SLICES=$sl timeout 200 python tools/dev_bench.py --one --iters=8 2>&1 | tail -1 >> $f
done
for pad in 14000 33000; do
  RMI_DEV_LEAF_SMEM_PAD=$pad timeout 200 python tools/dev_bench.py --one --iters=8 2>&1 | tail -1 >> $f
done
for hint in 00 21; do
  RMI_DEV_LEAF_SLICES=1 RMI_DEV_L2_HINT=$hint timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct --clock-control none \
     -k regex:k_leaf --launch-skip 4 --launch-count 2 --csv --log-file $out/${tag}_ncu_hint${hint}.csv python tools/dev_bench.py --one --iters=4 > /dev/null 2>&1
done
cat $f

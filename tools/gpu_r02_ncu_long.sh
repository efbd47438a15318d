#!/bin/bash
# ncu full captures of the leaf kernel on a build with 1525-key training vectors (what one GPU of an 8-GPU build sees)
tag=${1:-r02o}
out=gpurun_out
mkdir -p $out
for bf in ${2:-131072}; do
  RMI_DEV_LEAF_SLICES=1 timeout 120 python tools/dev_bench.py --spec=linear,linear --bf=$bf --iters=6 2>&1 | grep '^{' | cut -c1-400
  RMI_DEV_LEAF_SLICES=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_leaf -s 3 -c 1 \
     -o $out/${tag}_k_leaf_bf$bf -f python tools/dev_bench.py --spec=linear,linear --bf=$bf --iters=5 > $out/${tag}_ncu_$bf.log 2>&1
done
ls -la $out/*.ncu-rep

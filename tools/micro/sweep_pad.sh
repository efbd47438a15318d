for pad in 0 14000 33000 70000; do
  echo "pad $pad"
  RMI_DEV_LEAF_SMEM_PAD=$pad python tools/dev_bench.py --one --iters=4 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['phases_ms'])"
  RMI_DEV_LEAF_SMEM_PAD=$pad ncu --metrics dram__bytes_read.sum,gpu__time_duration.sum --clock-control none -k regex:k_leaf --launch-skip 3 --launch-count 1 python tools/dev_bench.py --one --iters=2 2>&1 | grep -E "dram__bytes_read|gpu__time"
done

#!/bin/bash
# Round 2, GPU call A: parity of the rewritten leaf kernel, A/B timing of library variants, ncu capture.
tag=${1:-r02a}
out=gpurun_out
mkdir -p $out
export PYTHONUNBUFFERED=1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $out/${tag}_smi.txt 2>&1
# 1. fast parity subset first (fails fast if the new kernel is wrong)
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_golden.py -x -q -m gpu > $out/${tag}_pytest_parity.log 2>&1
echo "pytest parity exit $?" >> $out/${tag}_pytest_parity.log
tail -3 $out/${tag}_pytest_parity.log
# 2. variants A/B (leaf phase of the headline build)
f=$out/${tag}_variants.jsonl
: > $f
for lib in rmi_b200/lib/librmi_b200*.so; do
  RMI_DEV_PRINT_OCC=1 RMI_B200_LIB=$PWD/$lib timeout 120 python tools/dev_bench.py --one --iters=8 2> $out/${tag}_occ_$(basename $lib .so).txt | tail -1 | sed "s#^{#{\"lib\": \"$(basename $lib)\", #" >> $f
  grep -h "k_leaf" $out/${tag}_occ_$(basename $lib .so).txt | sort -u | head -3
done
python - "$f" <<'PY'
import json, sys
rows = [json.loads(l) for l in open(sys.argv[1]) if l.startswith("{")]
for r in sorted(rows, key=lambda r: r.get("leaf_ms_min", 9e9)):
    if "leaf_ms_min" in r:
        print(f'{r["lib"]:34s} leaf {r["leaf_ms_min"]:.3f} ms  phases {[round(x,3) for x in r["phases_ms"]]} device {r["device_ms_min"]:.3f} ms  wall {r["wall_ms_min"]:.3f} ms')
    else:
        print(r)
PY
# 3. forcing fewer resident blocks on the default library (occupancy sensitivity)
for pad in 0 6000 16000; do
  RMI_DEV_LEAF_SMEM_PAD=$pad RMI_DEV_PRINT_OCC=1 timeout 120 python tools/dev_bench.py --one --iters=8 2>&1 | grep -h "leaf_ms_min\|blocks/SM" | sed "s/^{/{\"pad\": $pad, /" | cut -c1-400 >> $out/${tag}_pad.jsonl
done
tail -6 $out/${tag}_pad.jsonl | cut -c1-300
# 4. all dev configs on the default library
timeout 300 python tools/dev_bench.py --iters=3 > $out/${tag}_dev_bench.jsonl 2>&1
# 5. full GPU test suite
timeout 900 python -m pytest tests -m gpu -x -q > $out/${tag}_pytest_gpu.log 2>&1
echo "pytest exit $?" >> $out/${tag}_pytest_gpu.log
tail -3 $out/${tag}_pytest_gpu.log
# 6. bench + ncu
timeout 600 python bench.py > $out/${tag}_bench_n1.json 2> $out/${tag}_bench_n1.err
cat $out/${tag}_bench_n1.json | cut -c1-1500
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
   --log-file $out/${tag}_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $out/${tag}_ncu_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_leaf -s 10 -c 5 \
   -o $out/${tag}_k_leaf_full -f python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $out/${tag}_ncu_full.log 2>&1
# 7. memcheck of one small build
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 9 python __graft_entry__.py smoke > $out/${tag}_memcheck.log 2>&1
echo "memcheck exit $?" >> $out/${tag}_memcheck.log
tail -4 $out/${tag}_memcheck.log

#!/bin/bash
# Round 2, lean GPU call: fast parity subset, A/B timing of the library variants, one ncu capture, the contract bench.
tag=${1:-r02c}
out=gpurun_out
mkdir -p $out
export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_golden.py -x -q -m gpu > $out/${tag}_pytest_parity.log 2>&1
echo "pytest parity exit $?" >> $out/${tag}_pytest_parity.log
tail -3 $out/${tag}_pytest_parity.log
timeout 600 python -m pytest tests/test_gpu_sharded.py -x -q -m gpu -k "one_call or single_rank" > $out/${tag}_pytest_onecall.log 2>&1
echo "pytest one-call exit $?" >> $out/${tag}_pytest_onecall.log
tail -5 $out/${tag}_pytest_onecall.log
timeout 900 python -m pytest tests/test_optimizer.py tests/test_gpu_fullsize.py -x -q -m gpu -k "stats_batch or exact" > $out/${tag}_pytest_new.log 2>&1
echo "pytest new exit $?" >> $out/${tag}_pytest_new.log
tail -4 $out/${tag}_pytest_new.log
f=$out/${tag}_variants.jsonl
: > $f
for lib in rmi_b200/lib/librmi_b200*.so; do
  RMI_DEV_PRINT_OCC=1 RMI_B200_LIB=$PWD/$lib timeout 120 python tools/dev_bench.py --one --iters=8 2> $out/${tag}_occ_$(basename $lib .so).txt | tail -1 | sed "s#^{#{\"lib\": \"$(basename $lib)\", #" >> $f
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
timeout 300 python tools/dev_bench.py --iters=3 > $out/${tag}_dev_bench.jsonl 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_leaf -s 10 -c 3 \
   -o $out/${tag}_k_leaf_full -f python tools/dev_bench.py --one --iters=4 > $out/${tag}_ncu_full.log 2>&1
timeout 900 python bench.py > $out/${tag}_bench_n1.json 2> $out/${tag}_bench_n1.err
echo "bench exit $?"
cut -c1-3000 $out/${tag}_bench_n1.json
tail -5 $out/${tag}_bench_n1.err

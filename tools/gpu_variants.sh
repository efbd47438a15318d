#!/bin/bash
# On a GPU box: times the default library and every variant in rmi_b200/lib/ on the bench workload
# (tools/dev_bench.py --one), then runs the parity tests on the fastest variant.
#   gpurun -- 'bash tools/gpu_variants.sh r02a'
tag=${1:-variants}
out=gpurun_out
mkdir -p $out
f=$out/${tag}_variants.jsonl
: > $f
for lib in rmi_b200/lib/librmi_b200*.so; do
  RMI_B200_LIB=$PWD/$lib timeout 120 python tools/dev_bench.py --one --iters=8 2>&1 | tail -1 | sed "s#^{#{\"lib\": \"$(basename $lib)\", #" >> $f
done
python - "$f" <<'PY'
import json, sys
rows = [json.loads(l) for l in open(sys.argv[1]) if l.startswith("{")]
for r in sorted(rows, key=lambda r: r["leaf_ms_min"]):
    print(f'{r["lib"]:34s} leaf {r["leaf_ms_min"]:.3f} ms  device {r["device_ms_min"]:.3f} ms  wall {r["wall_ms_min"]:.3f} ms')
best = min(rows, key=lambda r: r["leaf_ms_min"])["lib"]
open(sys.argv[1] + ".best", "w").write(best)
PY
best=$(cat $f.best)
if [ "$best" != "librmi_b200.so" ]; then
  RMI_B200_LIB=$PWD/rmi_b200/lib/$best timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_golden.py -x -q -m gpu > $out/${tag}_parity_${best%.so}.log 2>&1
  tail -2 $out/${tag}_parity_${best%.so}.log
fi

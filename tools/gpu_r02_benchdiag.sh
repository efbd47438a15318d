#!/bin/bash
tag=${1:-r02l}
out=gpurun_out
mkdir -p $out
for mode in a b; do
  if [ $mode = b ]; then export RMI_BENCH_GC=off; fi
  RMI_BENCH_SAMPLER=none timeout 300 python bench.py --no-extras --no-cpu-baseline > $out/${tag}_bench_$mode.json 2> $out/${tag}_bench_$mode.err
  python - <<PY
import json
d=json.load(open("$out/${tag}_bench_$mode.json"))
print("$mode", "ms_per_step", round(d["ms_per_step"],3), "steps", d["details"]["step_wall_ms"])
PY
done

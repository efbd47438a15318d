#!/bin/bash
tag=${1:-r02i}
out=gpurun_out
mkdir -p $out
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_golden.py tests/test_gpu_fullsize.py -x -q -m gpu > $out/${tag}_pytest.log 2>&1
echo "pytest exit $?" >> $out/${tag}_pytest.log; tail -3 $out/${tag}_pytest.log
timeout 300 python tools/dev_bench.py --iters=8 --quick > $out/${tag}_dev_bench.jsonl 2>&1
python - <<PY
import json
for l in open("$out/${tag}_dev_bench.jsonl"):
    if l.startswith("{"):
        r=json.loads(l); print(r.get("spec"), r.get("bf"), [round(x,3) for x in r.get("phases_ms",[])], "dev", r.get("device_ms_min"), "wall min/med", round(r.get("wall_ms_min",0),3), round(r.get("wall_ms_med",0),3))
PY
timeout 900 python bench.py > $out/${tag}_bench_n1.json 2> $out/${tag}_bench_n1.err
echo "bench exit $?"; cut -c1-400 $out/${tag}_bench_n1.json

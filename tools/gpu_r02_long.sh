#!/bin/bash
# On a GPU box: A/B of the reciprocal look-ahead in the general fit step (librmi_b200.so vs librmi_b200_rc0.so) on
# builds with long training vectors, then the parity suites and the contract bench on the default library.
tag=${1:-r02n}
out=gpurun_out
mkdir -p $out
export PYTHONUNBUFFERED=1
for lib in $(cd rmi_b200/lib && ls librmi_b200*.so); do
  [ -f rmi_b200/lib/$lib ] || continue
  RMI_B200_LIB=$PWD/rmi_b200/lib/$lib timeout 300 python tools/dev_bench.py --long --iters=6 2>&1 | grep '^{' | sed "s#^{#{\"lib\": \"$lib\", #" >> $out/${tag}_long.jsonl
done
python - <<PY
import json
for l in open("$out/${tag}_long.jsonl"):
    r = json.loads(l)
    print(f'{r["lib"]:22s} {r.get("spec"):16s} {r.get("bf"):8d} leaf {r.get("leaf_ms_min",0):.3f} device {r.get("device_ms_min",0):.3f} wall {r.get("wall_ms_min",0):.3f}')
PY
if [ "$2" = quick ]; then
  timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_golden.py tests/test_gpu_fullsize.py -x -q -m gpu > $out/${tag}_pytest.log 2>&1
  echo "pytest exit $?" >> $out/${tag}_pytest.log; tail -3 $out/${tag}_pytest.log
  exit 0
fi
timeout 1500 python -m pytest tests -x -q -m gpu > $out/${tag}_pytest.log 2>&1
echo "pytest exit $?" >> $out/${tag}_pytest.log; tail -3 $out/${tag}_pytest.log
timeout 900 python bench.py > $out/${tag}_bench_n1.json 2> $out/${tag}_bench_n1.err
echo "bench exit $?"; cut -c1-600 $out/${tag}_bench_n1.json

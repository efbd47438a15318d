#!/bin/bash
# Final validation of a round: tools/gpu_round.sh plus the 4-ary boundary-search knob
# (parity subset + timing under RMI_DEV_BOUNDS_ARITY=4).
tag=${1:-final}
out=gpurun_out
mkdir -p $out
export PYTHONUNBUFFERED=1
RMI_DEV_BOUNDS_ARITY=4 timeout 200 python -m pytest tests/test_gpu_parity.py -x -q -m gpu > $out/${tag}_pytest_arity4.log 2>&1
echo "pytest arity4 exit $?" >> $out/${tag}_pytest_arity4.log
tail -2 $out/${tag}_pytest_arity4.log
: > $out/${tag}_arity.jsonl
for a in 2 4; do
  RMI_DEV_BOUNDS_ARITY=$a timeout 200 python tools/dev_bench.py --quick --iters=8 2>&1 | tail -3 >> $out/${tag}_arity.jsonl
done
cat $out/${tag}_arity.jsonl | cut -c1-400
bash tools/gpu_round.sh $tag ${2:-}

#!/bin/bash
# Final validation of a round: tools/gpu_round.sh (tests, smoke, bench, reference arm, ncu launch list).
bash tools/gpu_round.sh ${1:-final} ${2:-}

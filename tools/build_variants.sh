#!/bin/bash
set -o pipefail
# Builds experiment variants of librmi_b200.so next to the default library (CPU only: nvcc cross-compiles).
# Each variant is the same source with other compile-time knobs; tools/gpu_variants.sh times them on a GPU box.
#   tools/build_variants.sh            # all variants below
#   tools/build_variants.sh s3 t64     # a subset
declare -A DEFS=(
  [cf]="-DRMI_COOP_FORWARD=1"                       # warp-cooperative forward pass fed by 1-D bulk copies (cp.async.bulk + mbarrier tiles)
  [cfr]="-DRMI_COOP_FORWARD=1 -DRMI_FWD_BULK=0"     # the same fed by register look-ahead loads
  [rc0]="-DRMI_RC_PREFETCH=0"                       # general fit step loads 1/n in the step that uses it (no look-ahead)
  [lfa0]="-DRMI_LONG_FWD_ALL=0"                     # warps whose leaves are all long keep the lane-serial forward pass
  [lfa512]="-DRMI_LONG_FWD_MIN=512"                 # "long" from 512 keys on
  [ring0]="-DRMI_RCP_RING=0"                        # vectors past the shared table: general step (table / global table / division per item)
  [pu]="-DRMI_PARTIAL_UNROLLED=1"                   # chunks in which some lane ends: unrolled, predicated pieces instead of position-driven loops
  [s3]="-DRMI_SSTAGES=3 -DRMI_LEAF_MIN_BLOCKS=4"    # three copy stages in the ring (4 blocks per SM)
)
names=("$@")
[ ${#names[@]} -eq 0 ] && names=("${!DEFS[@]}")
pids=()
for n in "${names[@]}"; do
  echo "== variant $n: ${DEFS[$n]}"
  ( RMI_BUILD_TAG=$n RMI_NVCC_DEFS="${DEFS[$n]}" python rmi_b200/build.py > /tmp/build_variant_$n.log 2>&1 || { tail -20 /tmp/build_variant_$n.log; echo "variant $n FAILED"; exit 1; } ) &
  pids+=($!)
done
rc=0
for p in "${pids[@]}"; do wait $p || rc=1; done
[ $rc -ne 0 ] && exit 1
python rmi_b200/build.py > /tmp/build_default.log 2>&1 || { tail -20 /tmp/build_default.log; exit 1; }   # relink the CLI against the default library
ls -la rmi_b200/lib/

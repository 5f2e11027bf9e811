#!/bin/bash
# gpurun with retries while the pod answers "transient" / busy (nothing charged in that case).
#   scripts/gr.sh [--gpus N] <timeout_s> '<command>'
gp=""
if [ "$1" = "--gpus" ]; then gp="--gpus $2"; shift 2; fi
to=$1; shift
for i in $(seq 1 12); do
  out=$(/usr/local/graft/bin/gpurun $gp --timeout $to -- "$@" 2>&1)
  echo "$out" | tail -70
  if echo "$out" | grep -q "status=transient\|exit code 3\|status=busy"; then sleep 150; continue; fi
  break
done

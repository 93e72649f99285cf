#!/bin/bash
# usage: tools/gpurun_retry.sh <name> <timeout_s> [--gpus N] -- '<command>'
# Retries while the pod reports "no box / slot free right now" (exit code 3: nothing charged); stdout of the call in
# gpurun_out/<name>.stdout.
name=$1; shift
tmo=$1; shift
extra=()
while [ "$1" != "--" ]; do extra+=("$1"); shift; done
shift
for attempt in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout "$tmo" "${extra[@]}" -- "$1" > "gpurun_out/$name.stdout" 2>&1
  rc=$?
  if [ $rc -ne 3 ] && ! grep -q "status=transient" "gpurun_out/$name.stdout"; then
    echo "rc=$rc attempt=$attempt" >> "gpurun_out/$name.stdout"; exit $rc
  fi
  sleep 45
done
echo "gave up" >> "gpurun_out/$name.stdout"

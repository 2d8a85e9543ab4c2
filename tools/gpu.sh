#!/bin/bash
# gpurun with retries on "busy" (exit 3): tools/gpu.sh TIMEOUT 'command' [extra gpurun flags]
t=$1; shift; cmd=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout $t "$@" -- "$cmd"; rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 90
done
exit 3

#!/bin/bash
# gpurun with retries while the pod answers "busy" (exit 3); usage: tools/gpu_retry.sh [gpurun options] -- 'command'
for i in $(seq 1 12); do
  /usr/local/graft/bin/gpurun "$@"; rc=$?
  [ $rc -ne 3 ] && exit $rc
  sleep 60
done
exit 3

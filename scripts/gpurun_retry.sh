#!/bin/bash
# dev: retry a gpurun call while the pod answers "busy" (exit 3), up to 12 times 3 minutes apart
#   usage: scripts/gpurun_retry.sh [gpurun options] -- 'command'
for i in $(seq 1 12); do
  /usr/local/graft/bin/gpurun "$@"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 150
done
exit 3

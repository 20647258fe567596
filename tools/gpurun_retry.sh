#!/bin/bash
# retry gpurun while the pod answers "busy" (exit code 3); usage: gpurun_retry.sh <gpurun args...>
for i in $(seq 1 20); do
  /usr/local/graft/bin/gpurun "$@"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 120
done
exit 3

#!/bin/bash
# usage: tools/gpurun_retry.sh LOG [--gpus N] --timeout S -- 'cmd' ; retries while the pod answers busy (exit 3)
log=$1; shift
for i in $(seq 1 30); do
  /usr/local/graft/bin/gpurun "$@" > "$log" 2>&1; rc=$?
  if [ $rc -ne 3 ] && ! grep -q "status=transient" "$log"; then exit $rc; fi
  sleep 90
done
exit 3

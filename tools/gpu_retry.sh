#!/bin/bash
# Retries a gpurun call while the pod answers "transient" (exit 3: nothing charged).  usage: tools/gpu_retry.sh <timeout> '<command>'
for i in $(seq 1 12); do
    /usr/local/graft/bin/gpurun --timeout "$1" -- "$2"
    rc=$?
    if [ $rc -ne 3 ]; then exit $rc; fi
    sleep 150
done
exit 3

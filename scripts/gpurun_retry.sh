#!/bin/bash
# gpurun with retries while the pod answers "busy" (exit 3 / transient): usage gpurun_retry.sh LOG TIMEOUT [--gpus N] -- command
log=$1; shift
limit=$1; shift
for attempt in $(seq 1 40); do
    /usr/local/graft/bin/gpurun --timeout "$limit" "$@" > "$log" 2>&1
    rc=$?
    if [ $rc -ne 3 ] && ! grep -q "status=transient" "$log"; then echo "done rc=$rc attempt=$attempt" >> "$log"; exit $rc; fi
    sleep 90
done
echo "gave up" >> "$log"

#!/bin/bash
# usage: profiles/gpu_retry_n.sh <gpus> <tag> <timeout_s> <command...>   -- multi-GPU form of gpu_retry.sh
n=$1; shift; tag=$1; shift; to=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --gpus $n --timeout $to -- "$@" > gpurun_out/call_$tag.out 2>&1
  rc=$?
  if grep -q "status=transient" gpurun_out/call_$tag.out || [ $rc -eq 3 ]; then sleep 90; continue; fi
  break
done
echo "done rc=$rc try=$i" >> gpurun_out/call_$tag.out

#!/bin/bash
# Run each GPU test function in its own process (a device trap poisons the CUDA context of the
# process that hit it), each under a timeout; logs go to gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu_info.txt 2>&1
status=0
for t in "$@"; do
  name=$(echo "$t" | tr '/:' '__')
  echo "=== $t" | tee -a gpurun_out/isolated.log
  timeout -s KILL 420 python -m pytest -q -s -x -m gpu "$t" > "gpurun_out/$name.log" 2>&1
  rc=$?
  tail -n 40 "gpurun_out/$name.log" | grep -E "^\[gemm\]|passed|failed|error|Error|timeout|rel_err|assert" | tail -n 30 | tee -a gpurun_out/isolated.log
  echo "rc=$rc" | tee -a gpurun_out/isolated.log
  [ $rc -ne 0 ] && status=1
done
exit $status

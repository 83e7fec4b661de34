#!/bin/bash
# round 2, call Y: compute-sanitizer over smoke() (every product kernel on small inputs): memcheck, then racecheck and synccheck
mkdir -p gpurun_out
export CUDA_LAUNCH_BLOCKING=0
for tool in memcheck racecheck synccheck; do
  ( time timeout 900 compute-sanitizer --tool $tool --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/r2y_$tool.log 2>&1
  echo "== $tool"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|smoke c|Error|hazard" gpurun_out/r2y_$tool.log | head -12
done

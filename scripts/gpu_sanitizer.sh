#!/bin/bash
# compute-sanitizer: memcheck, racecheck, synccheck and initcheck over smoke() (every product kernel on small inputs, parity checked
# under the tool), then memcheck over the GPU test set (without the full-table whitelist and the reference-binding tests)
mkdir -p gpurun_out
export CUDA_LAUNCH_BLOCKING=0
for tool in memcheck racecheck synccheck; do
  ( time timeout 900 compute-sanitizer --tool $tool --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/sanitizer_$tool.log 2>&1
  echo "== $tool"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|smoke c|Error|hazard" gpurun_out/sanitizer_$tool.log | head -12
done
( time timeout 600 compute-sanitizer --tool initcheck --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/sanitizer_initcheck.log 2>&1
echo "== initcheck"; grep -E "ERROR SUMMARY|smoke c|Uninitialized" gpurun_out/sanitizer_initcheck.log | head -12
( time timeout 1500 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_whitelist_depth.py --deselect tests/test_gpu_binding.py ) > gpurun_out/sanitizer_memcheck_tests.log 2>&1
echo "== memcheck over tests"; grep -E "ERROR SUMMARY|passed|failed|Invalid" gpurun_out/sanitizer_memcheck_tests.log | head -12

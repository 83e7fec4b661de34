#!/bin/bash
# round 2, call Z: compute-sanitizer initcheck over smoke(), memcheck over the short GPU test set
mkdir -p gpurun_out
( time timeout 600 compute-sanitizer --tool initcheck --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/r2z_initcheck.log 2>&1
echo "== initcheck"; grep -E "ERROR SUMMARY|smoke c|Uninitialized" gpurun_out/r2z_initcheck.log | head -12
( time timeout 1500 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_whitelist_depth.py --deselect tests/test_gpu_binding.py ) > gpurun_out/r2z_memcheck_tests.log 2>&1
echo "== memcheck over tests"; grep -E "ERROR SUMMARY|passed|failed|Invalid" gpurun_out/r2z_memcheck_tests.log | head -12

#!/bin/bash
# what the driver runs at round end — all GPU tests, smoke, the reference arm and the default bench line
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q ) > gpurun_out/validate_tests.log 2>&1
tail -6 gpurun_out/validate_tests.log
( time python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/validate_smoke.log 2>&1
tail -6 gpurun_out/validate_smoke.log
( time timeout 600 python bench.py --impl reference --steps 10 --warmup 3 ) > gpurun_out/validate_bench_reference.json 2> gpurun_out/validate_bench_reference.err
( time timeout 1200 python bench.py ) > gpurun_out/validate_bench_default.json 2> gpurun_out/validate_bench_default.err
tail -c 600 gpurun_out/validate_bench_default.err
python scripts/bench_table.py gpurun_out/validate_bench_default.json | cut -c1-260
head -c 600 gpurun_out/validate_bench_reference.json

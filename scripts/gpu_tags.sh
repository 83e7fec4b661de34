mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tags.py tests/test_gpu_raw.py -x -q -m gpu 2>&1 | tail -25

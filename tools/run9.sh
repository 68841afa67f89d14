mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_hlgauss_gpu.py tests/test_twohot_gpu.py -m gpu -q -x 2>&1 | tail -25

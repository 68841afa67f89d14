mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_tc_gpu.py -m gpu -q -s -k "prefetched" 2>&1 | grep "PARITY\|passed\|failed\|Error" | tail -10

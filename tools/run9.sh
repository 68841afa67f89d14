mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_tc_gpu.py tests/test_update_shapes_gpu.py -m gpu -q -x 2>&1 | tail -4
timeout 900 python tools/trap_hunt.py 16 2>&1 | tee gpurun_out/r2_trap_hunt.txt | cut -c1-200 | tail -20
timeout 300 python tools/stream_bench.py 2>&1 | tee gpurun_out/r2_stream_bench2.jsonl | cut -c1-100

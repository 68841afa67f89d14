mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_tf32_gpu.py -m gpu -q --tb=short -s 2>&1 > gpurun_out/r2_pytest_tf32.log
grep -E "^E  |^FAILED|passed|failed|rel-L2|cosine" gpurun_out/r2_pytest_tf32.log | cut -c1-200 | head -40
timeout 300 python tools/tf32_bench.py sgemm 2>&1 | grep -v -i warn > gpurun_out/r2_tf32_bench.txt
cut -c1-175 gpurun_out/r2_tf32_bench.txt
timeout 300 python tools/tf32_bench.py gemm-only > gpurun_out/plain_tf32.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tf32_gemm_persist -s 3 -c 3 -f -o gpurun_out/r2_tf32_gemm_persist python tools/tf32_bench.py gemm-only > gpurun_out/r2_ncu_tf32.log 2>&1
tail -2 gpurun_out/r2_ncu_tf32.log

mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_tf32_gpu.py -m gpu -q --tb=short -s 2>&1 > gpurun_out/r2_pytest_tf32.log
grep -E "^E  |^FAILED|passed|failed|rel-L2" gpurun_out/r2_pytest_tf32.log | cut -c1-300 | head -60
: > gpurun_out/r2_tf32_bench.txt
for cfg in "1 4" "0 4" "2 4"; do set -- $cfg
MLB_TF32_ROUND=$1 MLB_TF32_STAGES=$2 timeout 300 python tools/tf32_bench.py 2>&1 | grep -v -i warn >> gpurun_out/r2_tf32_bench.txt
done
cut -c1-175 gpurun_out/r2_tf32_bench.txt

set -x
mkdir -p gpurun_out
for mode in 1 2; do
  export MLB_PARITY_LOG=$PWD/gpurun_out/r2_parity_epi$mode.jsonl; rm -f $MLB_PARITY_LOG
  MLB_TC_EPI=$mode timeout 900 python -m pytest tests/test_tc_gpu.py tests/test_update_shapes_gpu.py -m gpu -q -k "policy_forward_backward or update_iter" > gpurun_out/r2_pytest_epi$mode.log 2>&1; echo "pytest epi$mode rc=$?" >> gpurun_out/r2_pytest_epi$mode.log
  tail -4 gpurun_out/r2_pytest_epi$mode.log
done
rm -f gpurun_out/r2_epi_bench.jsonl
for mode in 0 1 2; do MLB_TC_EPI=$mode timeout 200 python tools/epi_bench.py >> gpurun_out/r2_epi_bench.jsonl 2>> gpurun_out/r2_epi_bench.err; done
cat gpurun_out/r2_epi_bench.jsonl
timeout 300 python -m pytest tests/test_phase1_gpu.py tests/test_golden_gpu.py -m gpu -q -x 2>&1 | tail -3
timeout 600 python tools/gae_sweep.py > gpurun_out/r2_gae_sweep2.log 2>&1; echo "sweep rc=$?"
grep "^T=" gpurun_out/r2_gae_sweep2.log | awk '$5>120'

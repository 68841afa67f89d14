mkdir -p gpurun_out
MLB_TC_EPI=1 timeout 600 python -m pytest tests/test_tc_gpu.py tests/test_update_shapes_gpu.py -m gpu -q -x -k "policy_forward_backward or update_iter" 2>&1 | tail -3
MLB_TC_EPI=1 timeout 120 ./tools/probe/phase_profile > gpurun_out/r2_phase_profile2.txt 2>&1; grep -A8 "dx_persist2" gpurun_out/r2_phase_profile2.txt
rm -f gpurun_out/r2_epi_bench3.jsonl
for mode in 0 1; do MLB_TC_EPI=$mode timeout 200 python tools/epi_bench.py >> gpurun_out/r2_epi_bench3.jsonl 2>> gpurun_out/r2_epi_bench.err; done
cat gpurun_out/r2_epi_bench3.jsonl
timeout 300 python tools/profile_update.py cfg2 2 > gpurun_out/r2_profile_cfg2.txt 2>&1; head -22 gpurun_out/r2_profile_cfg2.txt
timeout 300 python tools/profile_update.py cfg3 1 > gpurun_out/r2_profile_cfg3.txt 2>&1; head -22 gpurun_out/r2_profile_cfg3.txt

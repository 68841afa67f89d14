set -x
mkdir -p gpurun_out
MLB_TC_EPI=1 timeout 120 ./tools/probe/phase_profile > gpurun_out/r2_phase_profile.txt 2>&1; echo "rc=$?" >> gpurun_out/r2_phase_profile.txt
cat gpurun_out/r2_phase_profile.txt
rm -f gpurun_out/r2_epi_bench2.jsonl
for mode in 0 1; do MLB_TC_EPI=$mode timeout 200 python tools/epi_bench.py >> gpurun_out/r2_epi_bench2.jsonl 2>> gpurun_out/r2_epi_bench.err; done
cat gpurun_out/r2_epi_bench2.jsonl
for mode in 0 1; do
  MLB_TC_EPI=$mode timeout 600 python -m pytest tests/test_tc_gpu.py -m gpu -q -k "policy_forward_backward" 2>&1 | tail -2
done

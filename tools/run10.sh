mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"^dx_persist_kernel" -s 3 -c 1 -f -o gpurun_out/r2_dx_persist_v1 python tools/epi_bench.py > gpurun_out/r2_ncu_dx_persist_v1.log 2>&1; echo "ncu rc=$?"
MLB_PDL=0 timeout 300 python tools/profile_update.py cfg3 1 2>&1 | grep -v -i warn > gpurun_out/r2_profile_cfg3_stream.txt; head -14 gpurun_out/r2_profile_cfg3_stream.txt
MLB_PDL=0 timeout 300 python tools/profile_update.py cfg2 2 2>&1 | grep -v -i warn > gpurun_out/r2_profile_cfg2_final.txt; head -12 gpurun_out/r2_profile_cfg2_final.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"lstm_step_tc_kernel" -s 140 -c 1 -f -o gpurun_out/r2_lstm_step python tools/profile_update.py cfg4 1 > gpurun_out/r2_ncu_lstm_step.log 2>&1; echo "ncu rc=$?"

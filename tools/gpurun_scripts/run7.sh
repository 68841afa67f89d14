mkdir -p gpurun_out
MLB_TC_EPI=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:dx_persist2 -s 6 -c 1 -f -o gpurun_out/r2_dx2 python tools/epi_bench.py > gpurun_out/r2_ncu_dx2.log 2>&1; echo "ncu rc=$?"
MLB_TC_EPI=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:fwd_persist2 -s 6 -c 1 -f -o gpurun_out/r2_fwd2 python tools/epi_bench.py > gpurun_out/r2_ncu_fwd2.log 2>&1; echo "ncu rc=$?"
ls -la gpurun_out/*.ncu-rep

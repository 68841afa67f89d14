mkdir -p gpurun_out
timeout 300 python tools/tf32_bench.py gemm-only > gpurun_out/plain_tf32.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tf32_gemm -s 6 -c 3 -f -o gpurun_out/r2_tf32_gemm python tools/tf32_bench.py gemm-only > gpurun_out/r2_ncu_tf32.log 2>&1
tail -3 gpurun_out/r2_ncu_tf32.log
ls -la gpurun_out/r2_tf32_gemm.ncu-rep

mkdir -p gpurun_out
timeout 300 python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/plain_bench.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r2_launches_bench_final.csv python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
tail -1 gpurun_out/plain_bench.log | cut -c1-200
wc -l gpurun_out/r2_launches_bench_final.csv

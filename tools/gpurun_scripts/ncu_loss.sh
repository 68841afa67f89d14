mkdir -p gpurun_out
timeout 200 python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/plain_bench.log 2>&1 &&
timeout 400 ncu --set full --clock-control none --import-source on -k regex:ppo_loss_kernel -s 20 -c 2 -f -o gpurun_out/r2_ppo_loss_persist python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/ncu_loss.log 2>&1
tail -2 gpurun_out/ncu_loss.log

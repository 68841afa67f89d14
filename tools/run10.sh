mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_phase1_gpu.py -m gpu -q -k "dp_assign" 2>&1 | tail -3
for b in auto 64 32; do if [ $b = auto ]; then timeout 120 python tools/gae_block.py; else MLB_GAE_BLOCK=$b timeout 120 python tools/gae_block.py; fi; done 2>&1 | grep block | tee gpurun_out/r2_gae_block.jsonl
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"ppo_loss_kernel" -s 3 -c 1 -f -o gpurun_out/r2_ppo_loss python tools/profile_update.py cfg2 1 > gpurun_out/r2_ncu_ppo_loss.log 2>&1; echo "ncu rc=$?"

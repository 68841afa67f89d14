set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2_smi.txt 2>&1
./tools/probe/tmem_probe > gpurun_out/r2_tmem_probe.txt 2>&1; echo "probe rc=$?" >> gpurun_out/r2_tmem_probe.txt
export MLB_PARITY_LOG=$PWD/gpurun_out/r2_parity_measured.jsonl
rm -f $MLB_PARITY_LOG
timeout 1500 python -m pytest tests -m gpu -q -x --durations=15 > gpurun_out/r2_pytest1.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest1.log
tail -5 gpurun_out/r2_pytest1.log
timeout 900 python bench.py > gpurun_out/r2_bench1.json 2> gpurun_out/r2_bench1.err; echo "bench rc=$?"
tail -c 600 gpurun_out/r2_bench1.err
timeout 600 python tools/gae_sweep.py > gpurun_out/r2_gae_sweep1.log 2>&1; echo "sweep rc=$?"
tail -3 gpurun_out/r2_gae_sweep1.log

mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" 2>&1 | tail -2
timeout 1500 python -m pytest tests -m gpu -q --tb=short 2>&1 > gpurun_out/r2_pytest_full.log
grep -E "^E  |^FAILED|^ERROR|passed|failed" gpurun_out/r2_pytest_full.log | cut -c1-300 | head -20
timeout 300 python bench.py --no-extras --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
b=json.loads(sys.stdin.readline()); print(b['ms_per_step'], b['value'], b['e2e']['value'], b['gpu_launches'], b['clocks'])"

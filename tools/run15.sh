mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --tb=short 2>&1 > gpurun_out/r2_pytest_full.log
grep -E "^E  |^FAILED|^ERROR|passed|failed" gpurun_out/r2_pytest_full.log | cut -c1-300 | head -30
timeout 900 python bench.py > gpurun_out/r2_bench_final.json 2> gpurun_out/r2_bench_final.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench_final.json').read().strip().splitlines()[-1])
print('value', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'])
for k in ('f32','cfg3','cfg4'):
    print(k, d[k]['value'], d[k]['ms_per_step'])
PY
MLB_PREFETCH_GATHER=0 timeout 900 python bench.py --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
d=json.loads(sys.stdin.readline()); print('prefetch=0 value', d['value'])
for k in ('f32','cfg3','cfg4'): print(k, d[k]['value'], d[k]['ms_per_step'])"

mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" 2>&1 | tail -2
timeout 1500 python -m pytest tests -m gpu -q --tb=short 2>&1 > gpurun_out/r2_pytest_full.log
grep -E "^E  |^FAILED|^ERROR|passed|failed" gpurun_out/r2_pytest_full.log | cut -c1-300 | head -20
timeout 900 python bench.py > gpurun_out/r2_bench_final.json 2> gpurun_out/r2_bench_final.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench_final.json').read().strip().splitlines()[-1])
print('value', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['clocks'])
for k in ('f32','cfg3','cfg4'):
    print(k, d[k]['value'], d[k]['ms_per_step'])
print(json.dumps(d['roofline'])[:300])
for r in d['kernels']: print(r['kernel'][:40], r['us_per_launch'], round(r['share'],3))
PY

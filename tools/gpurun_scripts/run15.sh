for h in 0 1 0 1; do MLB_BWD_STREAMS=$h timeout 600 python bench.py --no-extras --no-cpu-baseline --steps 60 --warmup 20 2>/dev/null | python -c "
import sys, json
b=json.loads(sys.stdin.readline()); print('bwd_streams=$h', b['ms_per_step'], b['value'], b['e2e']['value'])"; done

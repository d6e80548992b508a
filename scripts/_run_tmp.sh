for t in 6 8 10 12; do python bench.py --no-cpu-baseline --e2e-threads $t 2>&1 | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['ms_per_step'],3), d['e2e'])"; done

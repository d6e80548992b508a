for i in 1 2 3 4 5 6 7 8 9 10; do SC2_TRANSFORM_PRIORITY=-1 python bench.py --no-cpu-baseline --no-e2e 2>&1 | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['ms_per_step'],3), d['clocks'])"; done

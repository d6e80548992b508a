for a in 4 4 4 4 8 8 8 8; do python bench.py --no-cpu-baseline --no-e2e --max-ahead $a 2>&1 | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('ahead $a', round(d['value']), round(d['ms_per_step'],3), 'issue', round(d['host_issue_ms_per_step'],3), 'mallocs', d['cudaMalloc_calls_in_timed_region'])"; done

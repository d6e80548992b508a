python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python bench.py --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','ms_per_step','one_batch_latency_ms')}); print(d['e2e'])
for k in d['kernels']:
    if k['stream']=='transform': print(k['kernel'], round(k['avg_launch_ms'],3))"

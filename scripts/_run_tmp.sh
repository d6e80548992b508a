for c in 8 32; do for i in 1 2 3 4 5; do CUDA_DEVICE_MAX_CONNECTIONS=$c python bench.py --no-cpu-baseline --no-e2e 2>&1 | tail -1 | python -c "
import json,sys,os; d=json.loads(sys.stdin.read()); print('conn', os.environ.get('CUDA_DEVICE_MAX_CONNECTIONS'), round(d['value']), round(d['ms_per_step'],3))"; done; done

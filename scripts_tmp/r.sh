timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/b9.json 2>/dev/null; python - <<'PY'
import json
d=json.loads(open('gpurun_out/b9.json').read().strip().splitlines()[-1])
k=d['roofline']['kernel_ms_per_step']
print(d['value'], d['ms_per_step'], {a:b for a,b in k.items() if 'token' in a or 'dpre' in a or 'fc1' in a})
PY

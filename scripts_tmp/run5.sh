timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench7.json 2> gpurun_out/bench7.err; tail -1 gpurun_out/bench7.json | cut -c1-200; tail -2 gpurun_out/bench7.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench7.json').read().strip().splitlines()[-1])
k=d['roofline']['kernel_ms_per_step']
print({a:b for a,b in k.items() if 'token' in a or 'slice' in a})
PY

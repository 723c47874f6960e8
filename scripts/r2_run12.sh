set -x
python -m pytest tests -m gpu -x -q --durations=6 > gpurun_out/r2_pytest12.log 2>&1; echo "pytest rc=$?"
python bench.py --steps 10 --warmup 3 --e2e-steps 3 --no-cpu-baseline --no-parity > gpurun_out/r2_bench12.json 2> gpurun_out/r2_bench12.err; echo "bench rc=$?"
tail -12 gpurun_out/r2_pytest12.log
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2_bench12.json").read().strip().splitlines()[-1])
k=d['roofline']['kernels']
print(round(d['value']), round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value']), {a:round(b['avg_ms'],3) for a,b in k.items() if b['avg_ms']>0.05})
for n,c in d.get('configs',{}).items(): print(n, round(c['value']), round(c['ms_per_step'],3), {a:b for a,b in (c.get('kernels_ms') or {}).items() if b>0.05})
PY

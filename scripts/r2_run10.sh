set -x
python bench.py --steps 10 --warmup 3 --e2e-steps 3 --no-cpu-baseline --no-parity > gpurun_out/r2_bench10.json 2> gpurun_out/r2_bench10.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2_bench10.json").read().strip().splitlines()[-1])
print(round(d['value']), round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value']))
for n,c in d.get('configs',{}).items(): print(n, round(c['value']), round(c['ms_per_step'],3), c.get('kernels_ms'))
PY

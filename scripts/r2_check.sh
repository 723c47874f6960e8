# quick confirmation of a build on one B200: GPU parity tests, then the default bench line (side configs included)
set -x
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/check_pytest.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/check_pytest.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/check_bench.json 2> gpurun_out/check_bench.err; echo "bench rc=$?"
tail -3 gpurun_out/check_bench.err
python - <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/check_bench.json").read().strip().splitlines()[-1])
    print(round(d['value']), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), 'parity', d.get('parity',{}).get('equal'), d.get('parity',{}).get('http_bodies_equal'))
    print({a:round(b['avg_ms'],3) for a,b in d['roofline']['kernels'].items()})
    for n,c in (d.get('configs') or {}).items(): print(n, round(c['value']), round(c['ms_per_step'],3), c.get('parity_equal'), c.get('kernels_ms'))
except Exception as e: print('ERR', e)
PY

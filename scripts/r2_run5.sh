set -x
python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest5.log 2>&1; echo "pytest rc=$?" 
python bench.py --steps 10 --warmup 3 --e2e-steps 3 --no-cpu-baseline --no-side-configs > gpurun_out/r2_bench5.json 2> gpurun_out/r2_bench5.err; echo "bench rc=$?"
tail -3 gpurun_out/r2_pytest5.log
tail -3 gpurun_out/r2_bench5.err
python - <<'PY'
import json
for f in ("gpurun_out/r2_bench5.json",):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        k=d['roofline']['kernels']
        print(f, round(d['value']), round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value']), {a:round(b['avg_ms'],3) for a,b in k.items()})
    except Exception as e: print(f, 'ERR', e)
PY

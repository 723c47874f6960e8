set -x
python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest9.log 2>&1; echo "pytest rc=$?"
python bench.py --steps 10 --warmup 3 --e2e-steps 3 --no-cpu-baseline > gpurun_out/r2_bench9.json 2> gpurun_out/r2_bench9.err; echo "bench rc=$?"
timeout 900 python scripts/mixed_probe.py 65536 > gpurun_out/r2_mixed_probe9.log 2>&1; echo "probe rc=$?"
tail -3 gpurun_out/r2_pytest9.log
python - <<'PY'
import json
for f in ("gpurun_out/r2_bench9.json",):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        k=d['roofline']['kernels']
        print(f, round(d['value']), round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value']), {a:round(b['avg_ms'],3) for a,b in k.items()})
        for n,c in d.get('configs',{}).items(): print(n, round(c['value']), round(c['ms_per_step'],3), c.get('parity_equal'))
    except Exception as e: print(f, 'ERR', e)
PY
cat gpurun_out/r2_mixed_probe9.log | tail -30

set -x
for k in 1 2 3; do python -m pytest tests/test_gpu_parity.py -x -q -k small_chunks > gpurun_out/r2_small$k.log 2>&1; echo "small_chunks run $k rc=$?"; done
timeout 1200 compute-sanitizer --tool initcheck python -m pytest tests/test_gpu_parity.py -x -q -k "small_chunks and not mixed and not full_size" > gpurun_out/r2_initcheck8.log 2>&1; echo "initcheck rc=$?"
timeout 600 compute-sanitizer --tool racecheck python scripts/repro_wkt.py small one > gpurun_out/r2_racecheck8.log 2>&1; echo "racecheck rc=$?"
python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest8.log 2>&1; echo "pytest rc=$?"
python bench.py --steps 10 --warmup 3 --e2e-steps 3 --no-cpu-baseline > gpurun_out/r2_bench8.json 2> gpurun_out/r2_bench8.err; echo "bench rc=$?"
tail -3 gpurun_out/r2_small1.log; tail -3 gpurun_out/r2_pytest8.log
grep -c "Uninitialized" gpurun_out/r2_initcheck8.log; grep -A14 "Uninitialized" gpurun_out/r2_initcheck8.log | head -80
grep -i "hazard\|ERROR SUMMARY" gpurun_out/r2_racecheck8.log | head
python - <<'PY'
import json
for f in ("gpurun_out/r2_bench8.json",):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        k=d['roofline']['kernels']
        print(f, round(d['value']), round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value']), {a:round(b['avg_ms'],3) for a,b in k.items()})
        for n,c in d.get('configs',{}).items(): print(n, round(c['value']), round(c['ms_per_step'],3), c.get('parity_equal'))
    except Exception as e: print(f, 'ERR', e)
PY

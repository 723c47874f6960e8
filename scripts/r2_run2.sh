set -x
python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest2.log 2>&1; echo "pytest rc=$?" 
python bench.py --steps 10 --warmup 3 --e2e-steps 3 > gpurun_out/r2_bench2.json 2> gpurun_out/r2_bench2.err; echo "bench rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2_launches2.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --e2e-steps 1 --no-side-configs --no-parity > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_encode_type|k_encode_tok3|k_encode_place" -s 3 -c 3 -o gpurun_out/r2_prof_walk2 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --e2e-steps 1 --no-side-configs --no-parity >/dev/null 2>&1
tail -5 gpurun_out/r2_pytest2.log
tail -3 gpurun_out/r2_bench2.err
python - <<'PY'
import json
for f in ("gpurun_out/r2_bench2.json",):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        k=d['roofline']['kernels']
        print(f, round(d['value']), round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value']), {a:round(b['avg_ms'],3) for a,b in k.items()})
        print('parity', d.get('parity_checked_items'), 'cpu', d.get('cpu_baseline'))
        print('configs', json.dumps(d.get('configs'))[:1500])
    except Exception as e: print(f, 'ERR', e)
PY

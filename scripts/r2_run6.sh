set -x
python -m pytest tests -m gpu -x -q --durations=5 > gpurun_out/r2_pytest6.log 2>&1; echo "pytest rc=$?" 
python bench.py --steps 10 --warmup 3 --e2e-steps 3 --no-cpu-baseline --no-side-configs > gpurun_out/r2_bench6.json 2> gpurun_out/r2_bench6.err; echo "bench rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_decode_coop_size|k_decode_coop_write|k_encode_tok3|k_encode_type" -s 6 -c 7 -o gpurun_out/r2_prof6 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --e2e-steps 1 --no-side-configs --no-parity >/dev/null 2>&1
tail -12 gpurun_out/r2_pytest6.log
tail -3 gpurun_out/r2_bench6.err
python - <<'PY'
import json
for f in ("gpurun_out/r2_bench6.json",):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        k=d['roofline']['kernels']
        print(f, round(d['value']), round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value']), {a:round(b['avg_ms'],3) for a,b in k.items()})
    except Exception as e: print(f, 'ERR', e)
PY

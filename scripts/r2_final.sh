# final evidence of the round on one B200: tests, the default bench line, the reference arm, a launch list and one
# `ncu --set full` capture of every lock-step kernel (read here afterwards: scripts/ncu_traffic.py, ncu_extract.py, ncu_lines.py)
set -x
python -m pytest tests -m gpu -x -q > gpurun_out/final_pytest.log 2>&1; echo "pytest rc=$?"
python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; echo "bench rc=$?"
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/final_bench_reference.json 2> gpurun_out/final_bench_reference.err; echo "reference rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/final_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --e2e-steps 1 --no-side-configs --no-parity > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_decode_coop_size|k_decode_coop_write|k_encode_tok3|k_encode_place|k_encode_type|k_encode_coop_emit" -s 9 -c 9 -o gpurun_out/final_prof python bench.py --steps 2 --warmup 1 --no-cpu-baseline --e2e-steps 1 --no-side-configs --no-parity > /dev/null 2>&1
tail -3 gpurun_out/final_pytest.log
tail -2 gpurun_out/final_bench.err
python - <<'PY'
import json
for f in ("gpurun_out/final_bench.json", "gpurun_out/final_bench_reference.json"):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d['value']), round(d['ms_per_step'],2), 'e2e', d.get('e2e'), d.get('cpu_baseline'))
        k=(d.get('roofline') or {}).get('kernels') or {}
        print({a:round(b['avg_ms'],3) for a,b in k.items()})
        for n,c in (d.get('configs') or {}).items(): print(n, round(c['value']), round(c['ms_per_step'],3), c.get('parity_equal'))
    except Exception as e: print(f, 'ERR', e)
PY

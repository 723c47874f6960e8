# last evidence of the round (the GPU tests ran green on this build minus the wait policy in scripts/r2_run22.sh): default bench
# line, reference arm, launch list, one `ncu --set full` capture of the lock-step kernels, then the GPU tests again
set -x
python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; echo "bench rc=$?"
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/final_bench_reference.json 2> gpurun_out/final_bench_reference.err; echo "reference rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/final_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --e2e-steps 1 --no-side-configs --no-parity > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_decode_coop_size|k_decode_coop_write|k_encode_tok3|k_encode_place|k_encode_type|k_encode_coop_emit" -s 9 -c 9 -o gpurun_out/final_prof -f python bench.py --steps 2 --warmup 1 --no-cpu-baseline --e2e-steps 1 --no-side-configs --no-parity > /dev/null 2>&1
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/final_pytest.log 2>&1; echo "pytest rc=$?"
tail -2 gpurun_out/final_pytest.log
python - <<'PY'
import json
for f in ("gpurun_out/final_bench.json", "gpurun_out/final_bench_reference.json"):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d['value']), round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value']), d['e2e'].get('host_wait'), (d.get('cpu_baseline') or {}).get('value'))
    except Exception as e: print(f, 'ERR', e)
PY

# last call of the round: GPU tests, a short bench line (value, e2e, whole-batch parity) and the `ncu --set full` capture of the
# lock-step kernels on the build with page-aligned payload copies
timeout 70 python -m pytest tests -m gpu -x -q > gpurun_out/final_pytest.log 2>&1; echo "pytest rc=$?"; tail -1 gpurun_out/final_pytest.log
timeout 55 python bench.py --steps 5 --warmup 3 --e2e-steps 5 --no-side-configs --no-cpu-baseline > gpurun_out/r28_bench.json 2> gpurun_out/r28_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/r28_bench.json").read().strip().splitlines()[-1])
    print(round(d['value']), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), 'bodies', round(d['e2e']['with_result_bodies'] or 0), 'parity', d.get('parity',{}).get('equal'), d['e2e'].get('host_wait'))
except Exception as e: print('ERR', e)
PY
timeout 80 ncu --set full --clock-control none --import-source on -k regex:"k_decode_coop_size|k_decode_coop_write|k_encode_tok3|k_encode_place|k_encode_type|k_encode_coop_emit" -s 9 -c 9 -o gpurun_out/final_prof -f python bench.py --steps 2 --warmup 1 --no-cpu-baseline --e2e-steps 1 --no-side-configs --no-parity > /dev/null 2>&1; echo "ncu rc=$?"

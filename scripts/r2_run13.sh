# routing threshold A/B (small items to the lock-step kernels), host pipeline trace and pipeline-shape sweep
set -x
export BENCH_ARGS="--no-side-configs --no-parity"
bash scripts/ab_bench.sh "GGR_X=0" "GGR_LOCKSTEP_MIN_BYTES=0" "GGR_LOCKSTEP_MIN_BYTES=256" > gpurun_out/r13_route_nested.log 2>&1
BENCH_ARGS="--no-side-configs --no-parity --workload flat" bash scripts/ab_bench.sh "GGR_X=0" "GGR_LOCKSTEP_MIN_BYTES=0" "GGR_LOCKSTEP_MIN_BYTES=128" > gpurun_out/r13_route_flat.log 2>&1
BENCH_ARGS="--no-side-configs --no-parity --workload mixed" bash scripts/ab_bench.sh "GGR_X=0" "GGR_LOCKSTEP_MIN_BYTES=0" > gpurun_out/r13_route_mixed.log 2>&1
GGR_TRACE=1 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 2 --no-side-configs --no-parity > gpurun_out/r13_trace.json 2> gpurun_out/r13_trace.log
for cfg in "GGR_SLOTS=4 GGR_CHUNK_ITEMS=8192" "GGR_SLOTS=4 GGR_CHUNK_ITEMS=4096" "GGR_SLOTS=6 GGR_CHUNK_ITEMS=4096" "GGR_SLOTS=8 GGR_CHUNK_ITEMS=2048" "GGR_SLOTS=3 GGR_CHUNK_ITEMS=16384" "GGR_SLOTS=4 GGR_CHUNK_ITEMS=12288"; do
env $cfg python bench.py --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 5 --no-side-configs --no-parity 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$cfg', 'value', round(d['value']), 'e2e', round(d['e2e']['value']), 'bodies', round(d['e2e']['with_result_bodies'] or 0))
"
done > gpurun_out/r13_sweep.log 2>&1
cat gpurun_out/r13_route_nested.log gpurun_out/r13_route_flat.log gpurun_out/r13_route_mixed.log gpurun_out/r13_sweep.log

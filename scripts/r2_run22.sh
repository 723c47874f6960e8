set -x
BENCH_ARGS="--no-side-configs --no-parity" bash scripts/ab_bench.sh "GGR_BLOCKING_SYNC=1" "GGR_BLOCKING_SYNC=0" "GGR_BLOCKING_SYNC=1" "GGR_BLOCKING_SYNC=0" 2>&1 | grep -v "^+" | cut -c1-120 > gpurun_out/r22_blocking.log
cat gpurun_out/r22_blocking.log
bash scripts/r2_final.sh

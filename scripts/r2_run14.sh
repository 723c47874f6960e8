# word-wise copies in the lock-step emit / write kernels, cached scan window, chunk ramp: tests, bench, pipeline-shape A/B
set -x
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/check_pytest.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/check_pytest.log
export BENCH_ARGS="--no-side-configs --no-parity"
bash scripts/ab_bench.sh "GGR_CHUNK_RAMP=1" "GGR_CHUNK_RAMP=0" "GGR_CHUNK_ITEMS=16384 GGR_SLOTS=3" "GGR_CHUNK_ITEMS=16384 GGR_SLOTS=4" "GGR_CHUNK_ITEMS=12288 GGR_SLOTS=4" 2>&1 | grep -v "^+" | cut -c1-700

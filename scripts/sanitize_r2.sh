# compute-sanitizer memcheck + racecheck over the small parity run on the last build of round 2 (bulk-copy tokenizer,
# word-wise copies); lock-step kernels forced for all item sizes
export GGR_LOCKSTEP_MIN_BYTES=0
for tool in memcheck racecheck; do
  echo "== $tool"
  timeout 150 compute-sanitizer --tool $tool --print-limit 5 python tests/gpu_sanitize_run.py > gpurun_out/sanitize_$tool.log 2>&1
  echo "rc=$?"
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|Race reported|hazard|Invalid|Barrier error|ok items|Error" gpurun_out/sanitize_$tool.log | sort | uniq -c | head -12
done

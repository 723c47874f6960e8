# compute-sanitizer passes over a small parity run (lock-step kernels forced for all item sizes)
export GGR_LOCKSTEP_MIN_BYTES=0
for tool in memcheck racecheck synccheck; do
  echo "== $tool"
  timeout 600 compute-sanitizer --tool $tool --print-limit 5 python tests/gpu_sanitize_run.py 2>&1 | grep -E "ERROR SUMMARY|Race reported|hazard|Invalid|Barrier error|ok items|Error" | sort | uniq -c | head -12
done

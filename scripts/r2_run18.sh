# first-tier table size of the reply size kernel (occupancy against overflow into the second tier); chunk size of the host pipeline with the ramp
set -x
bash scripts/ab_variants.sh base ent192 ent160 2>&1 | cut -c1-400
for cfg in "GGR_SLOTS=4 GGR_CHUNK_ITEMS=8192" "GGR_SLOTS=3 GGR_CHUNK_ITEMS=16384" "GGR_SLOTS=4 GGR_CHUNK_ITEMS=16384" "GGR_SLOTS=3 GGR_CHUNK_ITEMS=24576" "GGR_SLOTS=3 GGR_CHUNK_ITEMS=32768" "GGR_SLOTS=4 GGR_CHUNK_ITEMS=8192 GGR_CHUNK_RAMP=0"; do
env $cfg python bench.py --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 6 --no-side-configs --no-parity 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$cfg', 'value', round(d['value']), 'e2e', round(d['e2e']['value']), 'bodies', round(d['e2e']['with_result_bodies'] or 0))
"
done

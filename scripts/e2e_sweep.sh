# e2e (host-buffer calls) vs pipeline shape
for cfg in "GGR_SLOTS=8 GGR_CHUNK_ITEMS=8192" "GGR_SLOTS=8 GGR_CHUNK_ITEMS=4096" "GGR_SLOTS=6 GGR_CHUNK_ITEMS=6144" "GGR_SLOTS=5 GGR_CHUNK_ITEMS=10240"; do
env $cfg python bench.py --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 4 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$cfg', 'value', round(d['value']), 'e2e', round(d['e2e']['value']))
"
done

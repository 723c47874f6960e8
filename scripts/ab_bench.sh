# usage: scripts/ab_bench.sh "ENV1=.. ENV2=.." ["ENV..."]...   one short bench per environment setting
for envs in "$@"; do
env $envs python bench.py --steps 5 --warmup 3 --no-cpu-baseline --e2e-steps 2 ${BENCH_ARGS:-} 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
k=d['roofline']['kernels']
print('$envs', round(d['value']), round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value']), {a:round(b['avg_ms'],2) for a,b in k.items()})
"
done

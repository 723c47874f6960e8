# usage: scripts/bench_kernels.sh WORKLOAD ITEMS [extra bench args]: value, step and the per-kernel table, sorted
python bench.py --workload $1 --items $2 --steps 3 --warmup 3 --e2e-steps 1 --no-cpu-baseline --no-side-configs ${@:3} 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
k=d['roofline']['kernels']
print('$1', round(d['value']), 'transcodes/s', round(d['ms_per_step'],2), 'ms/step  e2e', round(d['e2e']['value']), 'parity', d.get('parity_checked_items'))
print('  ', ', '.join('%s %.3f' % (a, b['avg_ms']) for a,b in sorted(k.items(), key=lambda x:-x[1]['avg_ms']) if b['avg_ms'] > 0.02))
"

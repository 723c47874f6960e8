for n in 65536 75776 151552 303104; do
GGR_NO_COOP=1 python bench.py --steps 5 --warmup 3 --items $n --no-cpu-baseline --e2e-steps 2 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
k=d['roofline']['kernels']
print($n, round(d['value']), round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value']), {a:round(b['avg_ms'],2) for a,b in k.items()})
"
done

for v in base minb1 minb4 minb8; do
  if [ "$v" = base ]; then lib=""; else lib="$PWD/ggrmcp_b200/variants/libggrmcp_b200_$v.so"; fi
  for w in flat mixed; do
  GGR_LIB_PATH=$lib python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 1 --no-side-configs --no-parity 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
k=d['roofline']['kernels']
print('$v $w', round(d['value']), round(d['ms_per_step'],3), {a:round(b['avg_ms'],3) for a,b in k.items() if b['avg_ms']>0.05})
"
  done
done

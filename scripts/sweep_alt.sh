# A/B: k_encode_parse compiled for 4 / 6 / 8 resident blocks per SM, two batch sizes
for lib in libggrmcp_b200 libggr_alt_m6 libggr_alt_m8; do
for n in 113664 303104; do
GGR_LIB_PATH=$PWD/ggrmcp_b200/$lib.so GGR_NO_COOP=1 python bench.py --steps 5 --warmup 3 --items $n --no-cpu-baseline --e2e-steps 1 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
k=d['roofline']['kernels']
print('$lib', $n, round(d['value']), round(d['ms_per_step'],2), {a:round(b['avg_ms'],2) for a,b in k.items()})
"
done
done

# 2 GPUs of one box on the last build: the torchrun path the driver's scaling run takes
cat /sys/fs/cgroup/cpu.max
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29525 bench.py --gpus 2 --steps 5 --warmup 3 --e2e-steps 4 > gpurun_out/r25_bench2.json 2> gpurun_out/r25_bench2.err
echo "rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r25_bench2.json").read().strip().splitlines()[-1])
print(d['n_gpus'], round(d['value']), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), d['e2e'].get('per_rank'), d['e2e'].get('host_wait'), d['e2e'].get('pcie_probe_gbs_each_way_per_rank'))
PY

# 8 GPUs of one box: device-resident and end-to-end scaling (NUMA-local page-locked buffers, threads bound per GPU)
set -x
nvidia-smi topo -m > gpurun_out/r17_topo.txt 2>&1
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 8 --steps 5 --warmup 3 --e2e-steps 4 --no-cpu-baseline --no-side-configs --no-parity > gpurun_out/r17_bench8.json 2> gpurun_out/r17_bench8.err
echo "rc=$?"
tail -c 3000 gpurun_out/r17_bench8.json
tail -5 gpurun_out/r17_bench8.err

mkdir -p gpurun_out
nvidia-smi -L | head -4
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 4 --steps 10 --warmup 3 > gpurun_out/bench_n4.json 2> gpurun_out/bench_n4.err
echo "bench rc=$?"; grep '^{' gpurun_out/bench_n4.json | cut -c1-300; tail -3 gpurun_out/bench_n4.err

mkdir -p gpurun_out
nvidia-smi -L
( timeout 600 python -m pytest tests -m gpu -q -x -k "two_ranks or sharded" ) > gpurun_out/pytest_gpu_n2.log 2>&1; tail -4 gpurun_out/pytest_gpu_n2.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
echo "bench rc=$?"; cut -c1-400 gpurun_out/bench_n2.json; tail -5 gpurun_out/bench_n2.err

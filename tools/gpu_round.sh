#!/bin/bash
# one gpurun call: the -m gpu suite, the bench line, the launch list of the bench command (our kernels only), Cheetah decode timing
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q ${PYTEST_EXTRA:--x} ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
echo "bench rc=$?"; cut -c1-600 gpurun_out/bench_n1.json; tail -3 gpurun_out/bench_n1.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"^cham_|^prot_|^scan_g|^dec_|^chee_|^lion_|^cd_|decode_kernel|encode_kernel" --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-verify > /dev/null 2>&1
timeout 300 python tools/time_cheetah_decode.py 1024 text 2>&1 | tail -1 | tee gpurun_out/cheetah_decode.txt
ls -la gpurun_out | tail -8

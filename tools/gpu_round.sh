#!/bin/bash
# one gpurun call: the -m gpu suite, the bench line, the launch list of the bench command (our kernels only), ncu --set full of the two
# dominant kernels (DRAM traffic for bench.py's roofline.traffic)
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q ${PYTEST_EXTRA:--x} ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
echo "bench rc=$?"; cut -c1-400 gpurun_out/bench_n1.json; tail -3 gpurun_out/bench_n1.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"^cham_|^prot_|^scan_g|^dec_|^chee_|^lion_|^cd_|decode_kernel|encode_kernel" --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-verify > /dev/null 2>&1
if [ -n "$NCU_FULL" ]; then
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"^cham_flag_pass6" --launch-skip 5 --launch-count 1 -f -o gpurun_out/r2_flag6_final python tools/time_flag.py 1024 text 6 > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none -k regex:"^cham_emit" --launch-skip 1 --launch-count 1 -f -o gpurun_out/r2_emit_final python tools/time_flag.py 1024 text 6 > /dev/null 2>&1
fi
ls -la gpurun_out | tail -6

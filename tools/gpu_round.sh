#!/bin/bash
# One GPU call: full -m gpu suite, a short bench line, phase timing + ncu --set full of the flag pass, Cheetah decode timing + launch list.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/smi.txt 2>&1
( time timeout 1200 python -m pytest tests -m gpu -q ${PYTEST_EXTRA:--x} ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
echo "bench rc=$?"; cut -c1-1500 gpurun_out/bench_n1.json; tail -3 gpurun_out/bench_n1.err
DENSITY_B200_SO=density_b200/_variants/lib_timing.so timeout 300 python tools/time_phases.py > gpurun_out/phases.txt 2>&1; tail -4 gpurun_out/phases.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:^cham_flag_pass --launch-skip 5 --launch-count 1 -f -o gpurun_out/r2_flag python tools/time_phases.py > gpurun_out/ncu_flag.log 2>&1
timeout 300 python tools/time_cheetah_decode.py 1024 text > gpurun_out/cheetah_decode.txt 2>&1
timeout 300 python tools/time_cheetah_decode.py 256 mixed >> gpurun_out/cheetah_decode.txt 2>&1
cat gpurun_out/cheetah_decode.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"^cd_|^dec_|decode_tail" --csv --log-file gpurun_out/r2_cheetah_decode_launches.csv python tools/time_cheetah_decode.py 1024 text > /dev/null 2>&1
ls -la gpurun_out | tail -14

#!/bin/bash
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -q ${PYTEST_EXTRA:--x} ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
echo "bench rc=$?"; cut -c1-1200 gpurun_out/bench_n1.json; tail -3 gpurun_out/bench_n1.err
for k in 8 16 24; do
  DENSITY_B200_DEC_RUNS_PER_SM=$k timeout 300 python tools/time_cheetah_decode.py 1024 text 2>&1 | tail -1 | sed "s/^/runs_per_sm=$k /"
done | tee gpurun_out/cheetah_decode.txt
DENSITY_B200_DEC_RUNS_PER_SM=16 timeout 300 python tools/time_cheetah_decode.py 256 mixed 2>&1 | tail -1 | tee -a gpurun_out/cheetah_decode.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"^cd_|^dec_|decode_tail" --csv --log-file gpurun_out/r2_cheetah_decode_launches.csv python tools/time_cheetah_decode.py 1024 text > /dev/null 2>&1
ls -la gpurun_out | tail -8

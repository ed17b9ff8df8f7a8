mkdir -p gpurun_out
timeout 300 python tools/time_flag.py 256 random,mixed 6 --check 2>&1 | tee gpurun_out/time_flag.txt
timeout 300 python tools/time_flag.py 1024 mixed,text 6 --check 2>&1 | tee -a gpurun_out/time_flag.txt
timeout 300 python tools/bench_cheetah.py cheetah 256 --mixed --check 2>&1 | tail -3 | tee gpurun_out/chee_mixed.txt
( timeout 900 python -m pytest tests -m gpu -q -x -k "chameleon or cheetah or lion or paths" ) > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log

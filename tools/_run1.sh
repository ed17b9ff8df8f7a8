mkdir -p gpurun_out
timeout 300 python tools/time_flag.py 256 random,mixed 6 2>&1 | grep impl | tee gpurun_out/time_flag.txt
timeout 300 python tools/time_flag.py 1024 mixed,text 6 2>&1 | grep impl | tee -a gpurun_out/time_flag.txt
for a in cheetah lion; do timeout 300 python tools/bench_cheetah.py $a 1024 2>&1 | tail -1; timeout 300 python tools/bench_cheetah.py $a 256 --mixed --check 2>&1 | tail -2; done | tee gpurun_out/chee.txt
( timeout 1500 python -m pytest tests -m gpu -q -x -k "chameleon or cheetah_lion or paths or golden or kats" ) > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log

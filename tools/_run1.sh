mkdir -p gpurun_out
timeout 120 python tools/bench_cheetah.py cheetah 64 --check 2>&1 | tail -2 | tee gpurun_out/chee.txt
timeout 200 python tools/bench_cheetah.py cheetah 1024 --check 2>&1 | tail -2 | tee -a gpurun_out/chee.txt
timeout 200 python tools/bench_cheetah.py cheetah 1024 --p1 2>&1 | tail -1 | tee -a gpurun_out/chee.txt
timeout 200 python tools/bench_cheetah.py cheetah 256 --mixed --check 2>&1 | tail -2 | tee -a gpurun_out/chee.txt
( timeout 900 python -m pytest tests -m gpu -q -x -k "cheetah or kats or golden or sweep or instance" ) > gpurun_out/pytest_gpu.log 2>&1; tail -4 gpurun_out/pytest_gpu.log

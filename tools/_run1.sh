mkdir -p gpurun_out
timeout 300 python tools/time_flag.py 1024 text,mixed 1,6 --check 2>&1 | tee gpurun_out/time_flag.txt
timeout 200 python tools/time_flag.py 256 random 1,6 --check 2>&1 | tee -a gpurun_out/time_flag.txt
DENSITY_B200_SO=density_b200/_variants/lib_timing.so timeout 300 python tools/time_phases.py 2>&1 | tee gpurun_out/f6_phases.txt
( timeout 1200 python -m pytest tests -m gpu -q -x -k "chameleon or sharded or kats or golden or instance" ) > gpurun_out/pytest_gpu_cham.log 2>&1; tail -5 gpurun_out/pytest_gpu_cham.log

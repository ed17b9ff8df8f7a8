mkdir -p gpurun_out
timeout 300 python tools/time_flag.py 1024 text,mixed 6 --check 2>&1 | tee gpurun_out/time_flag.txt
timeout 200 python tools/time_flag.py 256 random 6 --check 2>&1 | tee -a gpurun_out/time_flag.txt

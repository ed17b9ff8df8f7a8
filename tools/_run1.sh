mkdir -p gpurun_out
timeout 300 python tools/time_decode.py 1024 text,mixed 1,7 2>&1 | tee gpurun_out/time_decode.txt
timeout 200 python tools/time_decode.py 256 random 1,7 2>&1 | tee -a gpurun_out/time_decode.txt
( timeout 1200 python -m pytest tests -m gpu -q -x -k "decode or kats or golden or instance or interchange" ) > gpurun_out/pytest_gpu_dec.log 2>&1; tail -4 gpurun_out/pytest_gpu_dec.log

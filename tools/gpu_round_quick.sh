#!/bin/bash
# the -m gpu suite and the bench line (no ncu)
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q -x ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
echo "bench rc=$?"; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_n1.json'))
print(d['value'], d['e2e']['value'], {k:v for k,v in d['extra'].items() if 'pageable' in k or 'pinned' in k})
PY
tail -3 gpurun_out/bench_n1.err

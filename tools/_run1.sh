mkdir -p gpurun_out
timeout 300 python tools/time_decode.py 1024 text 7 2>&1 | tee gpurun_out/time_decode.txt
( timeout 1200 python -m pytest tests -m gpu -q -x -k "decode or kats or golden or instance or interchange" ) > gpurun_out/pytest_gpu_dec.log 2>&1; tail -3 gpurun_out/pytest_gpu_dec.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"dec_|decode_pass" --csv --log-file gpurun_out/r2_decode_launches.csv python tools/time_decode.py 1024 text 7 > /dev/null 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r2_decode_launches.csv')) if len(r)>10 and r[0].isdigit()]
from collections import defaultdict
d=defaultdict(list)
for r in rows: d[r[4].split('(')[0][-50:]].append(float(r[-1])/1e6)
for k,v in d.items(): print(f"{k:52s} n={len(v):3d} max={max(v):.3f} ms")
PY

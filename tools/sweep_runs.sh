# sweep the runs-per-SM knob of the Cheetah / Lion encoders (results are bit-identical, only the speed changes)
for k in 4 8 12 16 24; do
  for alg in cheetah lion; do
    echo -n "runs/SM=$k  "; DENSITY_B200_RUNS_PER_SM=$k timeout 120 python tools/bench_cheetah.py $alg ${1:-1024} | tail -1
  done
done

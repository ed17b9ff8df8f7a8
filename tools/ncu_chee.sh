mkdir -p gpurun_out
for spec in "cheetah:chee_pass_p" "cheetah:chee_pass_c" "cheetah:chee_fold_p" "cheetah:chee_fold_c" "lion:lion_pass_p" "lion:lion_fold_p"; do
  alg=${spec%%:*}; k=${spec##*:}
  timeout 200 ncu --set full --clock-control none --import-source on -k regex:"^$k" --launch-skip 15 --launch-count 1 -f -o gpurun_out/r1_${k} python tools/bench_cheetah.py $alg 256 > /dev/null 2>&1
  ncu -i gpurun_out/r1_${k}.ncu-rep --page details --csv > gpurun_out/r1_${k}_details.csv 2>/dev/null
done
ls -la gpurun_out | tail -15

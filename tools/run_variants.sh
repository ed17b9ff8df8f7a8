for v in "" density_b200/_variants/lib_1024_8.so density_b200/_variants/lib_512_8.so density_b200/_variants/lib_256_16.so density_b200/_variants/lib_512_4.so; do
  echo "== variant: ${v:-default(1024x4)}"
  DENSITY_B200_SO=$v timeout 300 python -m pytest tests/test_gpu_parity.py -x -q --timeout 120 -m gpu -k "chameleon" 2>&1 | tail -1
  DENSITY_B200_SO=$v python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline | python -c "
import json,sys; d=json.loads(sys.stdin.read()); k=d['roofline']['kernels']; print('value %.1f GB/s  flag %.3f ms  mid %.3f  emit %.3f'%(d['value'], k['cham_flag_pass']['ms'], k['carry_resolve_sizes_scan']['ms'], k['cham_emit']['ms']))"
done

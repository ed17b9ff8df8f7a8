#!/bin/bash
# tools/build_variant.sh <name> "<extra nvcc flags>": builds density_b200/_variants/lib_<name>.so (run it with DENSITY_B200_SO=...)
set -e
name=$1; extra=$2
cd "$(dirname "$0")/../density_b200"
mkdir -p _variants _obj_v
objs=""
for f in api chameleon_encode chameleon_decode cheetah_encode cl_decode scalar_codec; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xcompiler -fvisibility=hidden --cudart static $extra -c csrc/$f.cu -o _obj_v/$f.o
  objs="$objs _obj_v/$f.o"
done
nvcc -shared -gencode arch=compute_100a,code=sm_100a --cudart static -o _variants/lib_$name.so $objs
echo built _variants/lib_$name.so

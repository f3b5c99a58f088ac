#!/bin/bash
# development: build a variant of the library with extra -D flags into build/var/<name>/libomega2d_b200.so
#   tools/build_variant.sh name -DO2D_BIG_TILE=1024 -DO2D_STAGES=2
set -e
cd "$(dirname "$0")/.."
name=$1; shift
out=build/var/$name; mkdir -p $out
NV="/usr/local/cuda/bin/nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -ccbin /usr/bin/g++ $*"
for f in o2d_api o2d_pp o2d_panels o2d_pp_rm o2d_pp_exp o2d_pp_wl o2d_pp_v2 o2d_pp_v3 o2d_substep o2d_bem o2d_reflect o2d_status o2d_vtu o2d_group; do
  $NV -c omega2d_b200/csrc/$f.cu -o $out/$f.o &
done
wait
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $out/libomega2d_b200.so $out/*.o -ldl -lpthread
rm -f $out/*.o
echo built $out/libomega2d_b200.so

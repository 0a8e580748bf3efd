#!/bin/bash
# Builds experimental copies of libpvp_b200.so with extra -D flags into exp_build/<name>.so (kernel experiments only;
# never shipped: the GPU script copies one over the in-tree library, benches, and restores the real one).
# usage: tools/build_exp.sh name "-DFLAG1 -DFLAG2"
set -e
cd "$(dirname "$0")/../pixeltovoxelprojector_b200/csrc"
mkdir -p ../../exp_build/obj_$1
for f in pvp_context pvp_motion pvp_image pvp_microbench pvp_pipeline pvp_analysis pvp_collective; do
  nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -fmad=false -prec-div=true -prec-sqrt=true -ftz=false -Xcompiler -fPIC,-O2,-ffp-contract=off $2 -c -o ../../exp_build/obj_$1/$f.o $f.cu &
done
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC,-O2 -x cu -c -o ../../exp_build/obj_$1/pvp_imageio.o pvp_imageio.cpp &
wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../../exp_build/$1.so ../../exp_build/obj_$1/*.o -lz
rm -rf ../../exp_build/obj_$1
echo built exp_build/$1.so

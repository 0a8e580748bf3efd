#!/bin/bash
# Builds an experimental copy of libpvp_b200.so with extra -D flags: pixeltovoxelprojector_b200/libpvp_b200_<name>.so, loaded with
# PVP_LIB_VARIANT=<name> (kernel A/B runs on one GPU box; never shipped, *.so is git-ignored).  Only the TU named by $3 (default
# pvp_motion) is recompiled with the flags; the other objects are those of the product build (make lib first).
# usage: tools/build_exp.sh name "-DFLAG1 -DFLAG2" [tu]
set -e
cd "$(dirname "$0")/../pixeltovoxelprojector_b200/csrc"
TU=${3:-pvp_motion}
mkdir -p build/exp
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -fmad=false -prec-div=true -prec-sqrt=true -ftz=false \
     -Xcompiler -fPIC,-O2,-ffp-contract=off -Xptxas -v $2 -c -o build/exp/$1_$TU.o $TU.cu 2> build/exp/$1_$TU.ptxas.log
OBJS=""
for f in pvp_context pvp_motion pvp_image pvp_microbench pvp_pipeline pvp_analysis pvp_collective pvp_imageio; do
  if [ $f = $TU ]; then OBJS="$OBJS build/exp/$1_$TU.o"; else OBJS="$OBJS build/$f.o"; fi
done
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../libpvp_b200_$1.so $OBJS -lz -Xlinker --exclude-libs,ALL
echo built pixeltovoxelprojector_b200/libpvp_b200_$1.so

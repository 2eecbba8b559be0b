#!/bin/bash
# build_variant.sh <name> <source.cu> "<extra nvcc flags>": build/librpgp_<name>.so = the current library with ONE source recompiled
# with extra -D flags (A/B experiments; tools/run_variants.sh swaps them in on the GPU box).  Runs in the build container (no GPU).
set -e
cd "$(dirname "$0")/.."
NAME=$1; SRC=$2; FLAGS=$3
CS=randomly-projected-additive-gps_b200/csrc
B=build/rpgp
make -s -j8 -C $CS > /dev/null
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xcompiler -O2 --expt-relaxed-constexpr $FLAGS -c $CS/$SRC -o build/variant_$NAME.o
OBJS=""
for o in $B/*.o; do
  if [ "$(basename $o .o)" != "$(basename $SRC .cu)" ]; then OBJS="$OBJS $o"; fi
done
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o build/librpgp_$NAME.so $OBJS build/variant_$NAME.o -cudart static
echo "built build/librpgp_$NAME.so"

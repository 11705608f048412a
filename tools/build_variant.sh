#!/bin/bash
# Build a codegen variant of the library for A/B runs: tools/build_variant.sh <name> "<extra nvcc flags>"
# -> tools/tune/lib_<name>.so (select it with SPECTRALMC_B200_LIB=$PWD/tools/tune/lib_<name>.so).
set -e
cd "$(dirname "$0")/../spectralmc_b200/csrc"
name=$1; shift
out=../../tools/tune/lib_${name}.so
tmp=$(mktemp -d)
for f in smc_api smc_normals smc_paths smc_cf smc_rowfft smc_cvnn smc_diag; do
  nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -Xptxas -v -I../../include "$@" -c $f.cu -o $tmp/$f.o 2> $tmp/$f.log &
done
for f in smc_normals smc_cf smc_diag; do  # the Philox4x32-7 build of the stream-drawing translation units
  nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -I../../include -DSMC_STREAM_P7 "$@" -c $f.cu -o $tmp/$f.p7.o 2> $tmp/$f.p7.log &
done
wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $out $tmp/*.o -lcudart
grep -A1 "tile_kernelIfLi0ELi0ELi0ELb0" $tmp/smc_cf.log | grep -E "registers|spill" | head -3
rm -rf $tmp
echo built $out

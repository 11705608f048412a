#!/bin/bash
# Build a codegen variant of smc_cf.cu only and link it with the in-tree objects of the other files:
#   tools/build_cf_variant.sh <name> "<extra nvcc flags>"   ->  tools/tune/lib_<name>.so
# (select it with SMC_LIB=... for tools/bench_raw.py or SPECTRALMC_B200_LIB=... for the package).
set -e
cd "$(dirname "$0")/../spectralmc_b200/csrc"
name=$1; shift
out=../../tools/tune/lib_${name}.so
tmp=$(mktemp -d)
nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -Xptxas -v -I../../include "$@" -c smc_cf.cu -o $tmp/smc_cf.o 2> $tmp/smc_cf.log &
nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -I../../include -DSMC_STREAM_P7 "$@" -c smc_cf.cu -o $tmp/smc_cf.p7.o 2> $tmp/smc_cf.p7.log &
wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $out smc_api.o smc_normals.o smc_normals.p7.o smc_paths.o $tmp/smc_cf.o $tmp/smc_cf.p7.o smc_rowfft.o smc_cvnn.o smc_diag.o smc_diag.p7.o -lcudart
echo "$name: $(grep -A2 '_ZN3smc11step_kernelIfLi0ELi0ELi0ELi0E' $tmp/smc_cf.log | grep -E 'registers|spill' | tr '\n' ' ')"
rm -rf $tmp

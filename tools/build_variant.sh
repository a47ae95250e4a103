#!/bin/bash
# usage: tools/build_variant.sh <name> "<extra nvcc flags>"   -> libeddsa_b200/variants/lib_<name>.so  (tuning only)
set -e
cd "$(dirname "$0")/../libeddsa_b200"
name=$1; extra=$2
d=variants/obj_$name; mkdir -p $d
for f in x25519 fixedbase verify; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -Xptxas -v -Xcompiler -fPIC -std=c++17 $extra -c csrc/kernels_$f.cu -o $d/kernels_$f.o 2> $d/kernels_$f.log &
done
wait
[ -f csrc/host.o ] || make csrc/host.o > /dev/null
nvcc -gencode arch=compute_100a,code=sm_100a -shared -cudart static -o variants/lib_$name.so $d/kernels_*.o csrc/host.o -Xlinker --version-script=csrc/exports.map -lpthread -ldl -lrt
echo "$name: $(grep -h -A1 "k_verify\|k_x25519E\|k_genpub\|k_signE" $d/*.log | grep -E "Used|spill" | sed 's/ptxas info    : //' | tr '\n' ' ' | cut -c1-400)"

#!/bin/bash
# builds A/B variants of libmeepo.so into meepoembedding_b200/ab/<name>.so:  scripts/build_ab.sh name "-DFLAG ..."
set -e
name=$1; flags=$2
cd "$(dirname "$0")/../meepoembedding_b200/csrc"
mkdir -p build_ab/$name ../ab
for f in *.cu; do
  /usr/local/cuda/bin/nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -ccbin /usr/bin/g++ -Xcompiler -fPIC,-fvisibility=hidden --fmad=false $flags -c $f -o build_ab/$name/${f%.cu}.o &
done
wait
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -ccbin /usr/bin/g++ -shared -o ../ab/$name.so build_ab/$name/*.o -cudart static -Xlinker -z,defs -ldl -lpthread -lrt
echo built ../ab/$name.so

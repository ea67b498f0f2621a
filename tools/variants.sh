#!/bin/bash
# builds tuning variants of libfovpt.so into variants/ (git-ignored); usage: tools/variants.sh name "-DX=1 -DY=2" ...
cd "$(dirname "$0")/../fovpathtracing_optixcodelatest_b200/csrc"
mkdir -p ../../variants
while [ $# -gt 1 ]; do
  name=$1; flags=$2; shift 2
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xcompiler -static-libstdc++ -Xcompiler -static-libgcc -Xlinker --exclude-libs,ALL $flags -shared -o ../../variants/lib_$name.so api.cu wavefront.cu bvh_build.cu -Xptxas -v 2> ../../variants/$name.log &
done
wait
grep -H -A3 "k_trace" ../../variants/*.log | grep Used

#!/bin/bash
# Development helper: build a variant of the library with extra -D flags for one source file.
#   tools/build_variant.sh NAME FILE.cu "-DFOO=1 ..."   ->  zstd_b200/variants/libzstd_b200_NAME.so
# Select it at run time with ZSTDB200_LIB=<path> (tests/wave_sweep.py, tests/profile_one.py).
set -e
cd "$(dirname "$0")/../zstd_b200/csrc"
name=$1; file=$2; flags=$3
mkdir -p ../variants /tmp/zbv_$name
make -s all
objs=""
for f in zb_api zb_dict zb_match zb_literals zb_sequences zb_stitch zb_decode; do
  if [ "$f.cu" == "$file" ]; then
    nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC,-fvisibility=hidden -Xptxas -v $flags -c $f.cu -o /tmp/zbv_$name/$f.o 2> /tmp/zbv_$name/$f.log
    grep "Used\|spill" /tmp/zbv_$name/$f.log | head -4 || true
    objs="$objs /tmp/zbv_$name/$f.o"
  else objs="$objs $f.o"; fi
done
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../variants/libzstd_b200_$name.so $objs -lcudart
echo built ../variants/libzstd_b200_$name.so

#!/bin/bash
# builds sda_b200/variants/lib_<name>.so: the library with extra nvcc flags for one source file (default packed_tc2.cu; A/B of
# kernel variants; pick one at run time with SDA_B200_LIB).  usage: tools/build_variant.sh name "-DSDA_TC2_XWIDE=1 ..." [file.cu]
set -e
name=$1; flags=$2; src=${3:-packed_tc2.cu}; obj=${src%.cu}.o
cd "$(dirname "$0")/../sda_b200/csrc"
make -s -j8
mkdir -p ../variants build/var_$name
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC,-fvisibility=hidden $flags -c $src -o build/var_$name/$obj
objs=$(ls build/*.o | grep -v "build/$obj")
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../variants/lib_$name.so $objs build/var_$name/$obj -Xlinker --exclude-libs,ALL
echo built sda_b200/variants/lib_$name.so

#!/bin/bash
# Builds tuning variants of liblqft.so into build/tune/.  Each argument: name:-DFLAG=..,-DFLAG=..
set -e
cd "$(dirname "$0")/.."
mkdir -p build/tune
for cfg in "$@"; do
  name=${cfg%%:*}; defs=${cfg#*:}; defs=${defs//,/ }
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -shared -I include \
    $defs -o build/tune/liblqft_$name.so lattice_qft_b200/csrc/lqft.cu -lcudart_static -ldl -lpthread -lrt &
done
wait
ls build/tune

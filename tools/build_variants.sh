#!/bin/bash
# Builds tuning variants of libjaadb200.so next to the product build (jaadec_b200/_build/variants/<name>.so); select one
# with JAADB200_LIB=<path>.  usage: tools/build_variants.sh name1 "-DFOO=1 -DBAR=2" name2 "..." ...
set -e
cd "$(dirname "$0")/../jaadec_b200/csrc"
mkdir -p ../_build/variants
while [ $# -ge 2 ]; do
  name=$1; defs=$2; shift 2
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 --fmad=false -Xcompiler -fPIC -shared $defs \
       -o ../_build/variants/$name.so jaadb_engine.cu container_index.cpp &
done
wait
ls -la ../_build/variants

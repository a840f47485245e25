#!/bin/bash
# Build a variant of libcdm_b200.so for kernel-tuning experiments: one translation unit is recompiled
# with extra -D flags, everything else is reused from the normal build.
#   scripts/build_variant.sh <tag> <file.cu> <extra nvcc flags...>
#   -> continuum-mechanics-mfem_b200/libcdm_b200_<tag>.so   (load it with CDM_B200_LIB=<path>)
set -e
tag=$1; src=$2; shift 2
cd "$(dirname "$0")/../continuum-mechanics-mfem_b200/csrc"
make -s -j8
mkdir -p build/var_$tag
base=${src%.cu}
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo \
   -Xcompiler -fPIC,-Wall,-Wno-unused-function -I../../include -I. "$@" -c $src -o build/var_$tag/$base.o
objs=""
for o in build/*.o; do
   if [ "$(basename $o)" = "$base.o" ]; then objs="$objs build/var_$tag/$base.o"; else objs="$objs $o"; fi
done
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../libcdm_b200_$tag.so $objs -ldl
echo "built continuum-mechanics-mfem_b200/libcdm_b200_$tag.so"

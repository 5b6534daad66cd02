#!/bin/bash
# usage: tools/build_variant.sh NAME [extra nvcc flags...]  ->  build/lib_NAME.so  (select it with $HB_LIB for A/B runs)
set -e
name=$1; shift
root=$(cd $(dirname $0)/.. && pwd)
src=${HB_SRC:-$root/huffman-gpu_b200/csrc}
out=$root/build/var_$name
mkdir -p $out
ARCH="-gencode arch=compute_100a,code=sm_100a"
for f in hb_api hb_encode hb_misc hb_comm hb_decode; do
  nvcc $ARCH -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xcompiler -fvisibility=hidden "$@" -I$src -c $src/$f.cu -o $out/$f.o &
done
gcc -O2 -std=c11 -fPIC -fvisibility=hidden -c $src/hb_codebook.c -o $out/hb_codebook.o
wait
nvcc $ARCH -shared -o $root/build/lib_$name.so $out/*.o -ldl
echo built build/lib_$name.so

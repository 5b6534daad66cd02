#!/bin/bash
# usage: tools/sass_dump.sh [G W C] -> /tmp/sass_gGwWcC.txt : the SASS of one encode_kernel<G, WIDE, CHECK> variant
G=${1:-4}; W=${2:-0}; C=${3:-1}
obj=$(dirname $0)/../huffman-gpu_b200/csrc/hb_encode.o
fn=$(cuobjdump -sass $obj | grep "Function :" | grep "encode_kernelILi${G}ELb${W}ELb${C}E" | awk '{print $3}')
cuobjdump -sass -fun "$fn" $obj | grep -E "^\s+/\*[0-9a-f]{4}\*/" | sed -E 's#^\s+/\*([0-9a-f]{4})\*/\s+#\1 #; s#\s*/\* 0x[0-9a-f]+ \*/##' > /tmp/sass_g${G}w${W}c${C}.txt
wc -l /tmp/sass_g${G}w${W}c${C}.txt

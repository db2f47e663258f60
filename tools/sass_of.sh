#!/bin/bash
# usage: tools/sass_of.sh <file.cu> <kernel-name-substring>  -> /tmp/<kernel>.sass (+ ptxas resource line)
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
SRC=$ROOT/mobrob_b200/csrc/$1
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -cubin -o /tmp/$1.cubin $SRC -Xptxas -v 2>&1 | grep -A2 "$2" | grep -E "registers|stack" || true
cuobjdump -sass /tmp/$1.cubin | awk -v k="$2" '/Function : /{f=index($0,k)>0} f' | grep -v "^\s*/\* 0x" | sed 's#/\* 0x[0-9a-f]* \*/##' | cut -c1-110 > /tmp/$2.sass
wc -l /tmp/$2.sass

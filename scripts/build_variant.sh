#!/bin/bash
# A-B builds: scripts/build_variant.sh <name> "<extra nvcc flags>" file.cu [file.cu ...]
# -> build/variants/<name>.so = the listed sources compiled with the extra flags + the regular objects of every other source.
# Use with CNS_LIB=build/variants/<name>.so (cns_b200/_lib.py).
set -e
cd "$(dirname "$0")/.."
NAME=$1; EXTRA=$2; shift 2
./build.sh > /dev/null
V=build/variants/$NAME.obj
mkdir -p $V
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC $EXTRA"
OBJS=""
for o in build/obj/*.o; do
  b=$(basename $o .o); use=$o
  for f in "$@"; do if [ "$(basename $f .cu)" == "$b" ]; then nvcc $FLAGS -c $f -o $V/$b.o; use=$V/$b.o; fi; done
  OBJS="$OBJS $use"
done
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o build/variants/$NAME.so $OBJS -lcudart
echo "built build/variants/$NAME.so"

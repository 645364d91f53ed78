#!/bin/bash
# Builds libcns_sm100.so in-tree (cross-compiles for sm_100a; no GPU needed).
# One object per .cu (compiled in parallel, rebuilt only when the source or any header is newer), then one link.
set -e
cd "$(dirname "$0")"
OUT=cns_b200/libcns_sm100.so
OBJ=build/obj
mkdir -p $OBJ
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC ${NVCC_EXTRA}"
NEWEST_HDR=$(ls -t cns_b200/csrc/*.cuh include/*.h build.sh | head -1)
TODO=""
for f in cns_b200/csrc/*.cu; do
  o=$OBJ/$(basename $f .cu).o
  if [ ! -f $o ] || [ $f -nt $o ] || [ $NEWEST_HDR -nt $o ]; then TODO="$TODO $f"; fi
done
if [ -n "$TODO" ]; then
  echo $TODO | tr ' ' '\n' | xargs -P "$(nproc)" -I{} sh -c "nvcc $FLAGS -c {} -o $OBJ/\$(basename {} .cu).o"
fi
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $OUT $OBJ/*.o -lcudart
echo "built $OUT ($(echo $TODO | wc -w) objects recompiled)"

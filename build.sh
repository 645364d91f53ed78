#!/bin/bash
# Builds libcns_sm100.so in-tree (cross-compiles for sm_100a; no GPU needed).
set -e
cd "$(dirname "$0")"
OUT=cns_b200/libcns_sm100.so
SRCS="cns_b200/csrc/handle.cu cns_b200/csrc/prep.cu cns_b200/csrc/encoder.cu cns_b200/csrc/encoder_tc.cu cns_b200/csrc/gemm_tc.cu cns_b200/csrc/backbone.cu cns_b200/csrc/dense_l1.cu cns_b200/csrc/backbone_fused.cu cns_b200/csrc/forward.cu"
for f in cns_b200/csrc/graph_build.cu cns_b200/csrc/backward.cu cns_b200/csrc/backward_kernels.cu cns_b200/csrc/affinity.cu cns_b200/csrc/train_ops.cu cns_b200/csrc/servo_edges.cu cns_b200/csrc/gemm_split.cu; do [ -f $f ] && SRCS="$SRCS $f"; done
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 \
     -Xcompiler -fPIC -shared ${NVCC_EXTRA} -o $OUT $SRCS -lcudart
echo "built $OUT"

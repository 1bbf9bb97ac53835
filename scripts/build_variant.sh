#!/bin/bash
# Build an A/B variant of the library next to the product one: build_variant.sh <name> <extra nvcc flags...>
# -> dis-yolo_b200/libdisyolo_b200_<name>.so ; run with DISYOLO_LIB=<that path>.
set -e
name=$1; shift
cd "$(dirname "$0")/../dis-yolo_b200/csrc"
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC --expt-relaxed-constexpr -shared "$@" \
  -o ../libdisyolo_b200_${name}.so net.cu conv_tc.cu conv_misc.cu postproc.cu train.cu train_tc.cu imgproc.cu augment.cu -lcudart_static -ldl -lrt -lpthread
echo built ../libdisyolo_b200_${name}.so

#!/bin/bash
# usage: build_variant_tc32.sh <name> <extra nvcc flags...>   -> yolo-fpga-accelerator_b200/lib/variants/libyolo2cuda_<name>.so
# (csrc/conv_i16_tc32.cu rebuilt with the flags, every other object from the product build)
set -e
name=$1; shift
cd /root/repo/yolo-fpga-accelerator_b200
mkdir -p lib/variants
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC "$@" -c -o lib/variants/tc32_$name.o csrc/conv_i16_tc32.cu
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o lib/variants/libyolo2cuda_$name.so lib/conv_i16.o lib/conv_i16_g1.o lib/conv_i16_tc2.o lib/variants/tc32_$name.o lib/conv_f32.o lib/bw_ops.o lib/capi.o lib/detect.o -cudart static
rm lib/variants/tc32_$name.o
echo built $name

#!/bin/bash
# tuning helper: build ../variants/lib_<name>.so with extra nvcc flags applied to sdv_kernels.cu only
# (A/B of tracker-kernel variants on the GPU box; select one with SDV_B200_LIB=<path>)
set -e
name=$1; shift
cd "$(dirname "$0")"
mkdir -p ../variants
/usr/local/cuda/bin/nvcc -O3 -std=c++17 -lineinfo --fmad=false -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC,-ffp-contract=off -Xptxas -v "$@" -c -o ../variants/kern_$name.o sdv_kernels.cu 2> ../variants/kern_$name.ptxas.log
/usr/local/cuda/bin/nvcc -shared -gencode arch=compute_100a,code=sm_100a -o ../variants/lib_$name.so ../variants/kern_$name.o sdv_capi.o sdv_ba.o sdv_ba_kernels.o sdv_refine.o sdv_reproject.o sdv_policy.o sdv_trace.o sdv_select.o sdv_lidar.o -cudart static
grep -A2 "track_cluster_kernelILi128" ../variants/kern_$name.ptxas.log | grep -E "registers|spill" | head -3

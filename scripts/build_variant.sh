#!/bin/bash
# tuning: build deal-and-ceed-on-gpu_b200/libbp5b200_<name>.so with extra nvcc flags (the .so travels to the GPU box;
# select it with BP5_LIB=...).  usage: scripts/build_variant.sh <name> "<flags>"
set -e
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
name=$1; flags=$2
W=/tmp/variant_$name
rm -rf $W; mkdir -p $W/deal-and-ceed-on-gpu_b200 $W/include
cp -r $ROOT/deal-and-ceed-on-gpu_b200/csrc $W/deal-and-ceed-on-gpu_b200/
cp -r $ROOT/include/* $W/include/
cd $W/deal-and-ceed-on-gpu_b200/csrc
rm -f *.o
make -j8 EXTRA="$flags" 2>&1 | grep -E " error |ld returned" || true
cp $W/deal-and-ceed-on-gpu_b200/libbp5b200.so $ROOT/deal-and-ceed-on-gpu_b200/libbp5b200_$name.so
echo built libbp5b200_$name.so

#!/bin/bash
# Build a kernel-variant library for A/B runs on the GPU box (select it with VSSEG_LIB_PATH):
#   tools/build_variant.sh NAME -DVSSEG_GATE_BATCH=4 ...   ->  vs_seg_b200/variants/libvsseg_b200_NAME.so
# Only vsseg_tc.cu takes the extra flags; the other objects come from the regular in-tree build.
set -e
name=$1; shift
cd "$(dirname "$0")/../vs_seg_b200/csrc"
make -s >/dev/null
mkdir -p ../variants
nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC "$@" -c vsseg_tc.cu -o ../variants/vsseg_tc_$name.o
objs=$(ls *.o | grep -v '^vsseg_tc\.o$')
nvcc -shared -gencode arch=compute_100a,code=sm_100a -o ../variants/libvsseg_b200_$name.so $objs ../variants/vsseg_tc_$name.o -lcudart
rm -f ../variants/vsseg_tc_$name.o
echo "built vs_seg_b200/variants/libvsseg_b200_$name.so"

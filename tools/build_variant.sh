#!/bin/bash
# usage: tools/build_variant.sh NAME "-DVARW_ATOMIC=0 ..."  -> build/variants/lib_NAME.so (experiment builds of mor_engine.cu)
set -e
cd "$(dirname "$0")/../lakesoul_b200/csrc"
NAME=$1; shift
mkdir -p ../../build/variants
PA=$(python -c "import pyarrow,os;print(os.path.dirname(pyarrow.__file__))")
/usr/local/cuda/bin/nvcc -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC,-fvisibility=hidden -I. -I../../include $@ -c mor_engine.cu -o ../../build/variants/me_$NAME.o
OBJ=../../build/obj
g++ -o ../../build/variants/lib_$NAME.so ../../build/variants/me_$NAME.o $OBJ/filter_kernels.cu.o $OBJ/decode_kernels.cu.o $OBJ/c_api.cc.o $OBJ/reader.cc.o $OBJ/writer.cc.o $OBJ/device_scan.cc.o $OBJ/filter.cc.o $OBJ/filter_eval.cc.o $OBJ/page_plan.cc.o \
  -shared -L/usr/local/cuda/lib64 -lcudart_static -ldl -lrt -lpthread -L$PA -l:libarrow.so.2400 -l:libparquet.so.2400 -Wl,-rpath,$PA -Wl,--no-undefined -Wl,--version-script=exports.map
rm -f ../../build/variants/me_$NAME.o
cuobjdump -res-usage ../../build/variants/lib_$NAME.so 2>/dev/null | grep -A1 "16var_write_kernelILi4" | grep REG

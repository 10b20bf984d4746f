# build_variant.sh NAME "FLAGS": libaec.so.0 with aec_skim.cu compiled with extra -D flags -> libaec_b200/lib/variants/NAME/libaec.so.0
set -e
NAME=$1; FLAGS=$2
D=libaec_b200/lib/variants/$NAME; mkdir -p $D
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC --use_fast_math $FLAGS -c libaec_b200/csrc/aec_skim.cu -o $D/aec_skim.o
OBJ=libaec_b200/lib/obj
nvcc -gencode arch=compute_100a,code=sm_100a -shared -Xcompiler -fPIC -Xlinker -soname,libaec.so.0 -o $D/libaec.so.0 $OBJ/aec_encode.o $OBJ/aec_decode.o $D/aec_skim.o $OBJ/aec_sz.o $OBJ/aec_runtime.o $OBJ/libaec_api.o $OBJ/sz_batch.o -lpthread
echo built $D/libaec.so.0

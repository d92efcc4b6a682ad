#!/bin/bash
# debugging aid: alternative builds of the G2 bucket-reduction translation unit, linked into separate libraries
#   v3: Fp2 base multiplications inlined (no calls)      v4: bucket add/double as calls as well      v5: ptxas -O1
cd "$(dirname "$0")/../algebra_b200/csrc" || exit 1
F="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC"
mkdir -p build_var
OTHERS=$(ls build/*.o | grep -v msm_inst_red_g2.o)
nvcc $F -DAB_FP2_INLINE_MUL -c msm_inst_red_g2.cu -o build_var/red_g2_v3.o &
nvcc $F -DAB_EC_NOINLINE_WIDE -c msm_inst_red_g2.cu -o build_var/red_g2_v4.o &
nvcc $F -Xptxas -O1 -c msm_inst_red_g2.cu -o build_var/red_g2_v5.o &
wait
for v in 3 4 5; do nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../libalgebra_b200_g2v$v.so $OTHERS build_var/red_g2_v$v.o -lcudart; done
ls -la ../*.so

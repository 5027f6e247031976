#!/bin/bash
# build a variant of the library with extra -D flags into /tmp-like path inside repo and run quick bench: usage variant.sh "<flags>"
cd aeonflux_b200/csrc && nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC,-fvisibility=hidden -shared $1 -o libaeonflux_b200.so afx_b200.cu

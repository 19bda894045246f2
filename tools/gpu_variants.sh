# A/B of experimental builds of the pair kernel at N = 1e6 (run through gpurun):
#   VARIANTS="a b" bash tools/gpu_variants.sh   compares the product library with build/variants/lib_a.so, lib_b.so
# A variant is the same library with pairbin.cu compiled under an experiment macro, e.g.
#   nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -Iinclude -DPB_EXP_X \
#        -c treegp_b200/csrc/pairbin.cu -o build/variants/pairbin_x.o
#   nvcc -gencode arch=compute_100a,code=sm_100a -shared -o build/variants/lib_x.so build/variants/pairbin_x.o \
#        build/kmat.o build/dense.o build/predict.o build/microbench.o build/hostrng.o build/vcorr.o
# (build/ travels to the GPU box; TREEGP_B200_LIB selects the library, see treegp_b200/_cabi.py).
cd /root/repo
for v in default ${VARIANTS:-}; do
  echo "VARIANT=$v"
  if [ "$v" != default ]; then export TREEGP_B200_LIB=/root/repo/build/variants/lib_$v.so; else unset TREEGP_B200_LIB; fi
  PB_N=1000000 PB_REPS=${REPS:-4} python tools/pb_run.py 2>&1 | tail -4
done

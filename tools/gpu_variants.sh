# A/B of experimental builds of the pair kernel (build/variants/lib_<name>.so; "default" = the product library), N = 1e6
cd /root/repo
for v in default ${VARIANTS:-nothr noboxpf neither}; do
  echo "VARIANT=$v"
  if [ "$v" != default ]; then export TREEGP_B200_LIB=/root/repo/build/variants/lib_$v.so; else unset TREEGP_B200_LIB; fi
  PB_N=1000000 PB_REPS=${REPS:-4} python tools/pb_run.py 2>&1 | tail -4
done

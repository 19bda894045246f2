# A/B of experimental builds of the pair kernel (build/variants/lib_*.so), N = 1e6, general dispatch vs short cuts
cd /root/repo
for v in "" bisect noguess both $EXTRA_VARIANTS; do
  for f in 0 3; do
    echo "VARIANT=${v:-default} FAST=$f"
    if [ -n "$v" ]; then export TREEGP_B200_LIB=/root/repo/build/variants/lib_$v.so; else unset TREEGP_B200_LIB; fi
    PB_FAST=$f PB_N=1000000 PB_REPS=3 python tools/pb_run.py 2>&1 | tail -3
  done
done

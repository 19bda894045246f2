# ncu --set full captures of the non-pairbin kernels the north star names, one GPU.
# Each probe is first run plain (must exit 0), then once under ncu for ONE launch of the named kernel.
cd /root/repo
TAG=${TAG:-r2a}
O=gpurun_out
cap() {  # name, kernel regex, skip, env..., -- command
  local name=$1 rx=$2 skip=$3; shift 3
  env "$@" > $O/plain_${name}_$TAG.log 2>&1 || { echo "plain run of $name failed"; tail -3 $O/plain_${name}_$TAG.log; return; }
  tail -4 $O/plain_${name}_$TAG.log
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$rx -s $skip -c 1 \
      -o $O/prof_${name}_$TAG -f env "$@" > $O/ncu_${name}_$TAG.log 2>&1 || echo "ncu of $name failed"
}
cap kmat_rbf      kmat_sym_kernel            1   PN=40000 python tools/kmat_probe.py
cap kmat_vk       kmat_sym_kernel            13  PN=40000 python tools/kmat_probe.py
cap kmat_cross_vk kmat_cross_kernel          7   PN=40000 python tools/kmat_probe.py
cap predict_trunc predict_mean_trunc_kernel  1   python tools/predict_probe.py
cap predict_full  "predict_mean_kernel"      1   PM=200000 python tools/predict_probe.py
cap trsv          trsv_step_kernel           200 DP_N=20000 python tools/dense_probe.py
cap panel_left    panel_left_kernel          100 DP_N=20000 python tools/dense_probe.py
cap vcorr         vcorr_kernel               1   python tools/vcorr_probe.py
cap pairbin_w     pairbin_kernel             1   PB_W=1 PB_N=1000000 PB_REPS=2 python tools/pb_run.py
ls -la $O/*.ncu-rep | tail -12

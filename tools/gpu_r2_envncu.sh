# ncu --set full captures inside one envelope likelihood evaluation (N = 40k): one trailing update (gemm_nt_sub_kernel)
# and one panel launch (panel_left_kernel) from the middle of the factorisation.
cd /root/repo
TAG=${TAG:-r2n}
O=gpurun_out
TRSV_NOCLUSTER=1 timeout 120 python tools/envelope_once.py || { echo "plain run failed"; exit 1; }
TRSV_NOCLUSTER=1 timeout 200 ncu --set full --clock-control none --import-source on -k regex:gemm_nt_sub_kernel -s 40 -c 1 \
    -o $O/prof_env_gemm_$TAG -f python tools/envelope_once.py > $O/ncu_env_gemm_$TAG.log 2>&1 || echo "gemm capture failed"
TRSV_NOCLUSTER=1 timeout 200 ncu --set full --clock-control none --import-source on -k regex:panel_left_kernel -s 324 -c 1 \
    -o $O/prof_env_panel_$TAG -f python tools/envelope_once.py > $O/ncu_env_panel_$TAG.log 2>&1 || echo "panel capture failed"
ls -la $O/prof_env_*_$TAG.ncu-rep

# Round-2 records of the shared-geometry bootstrap kernel (one GPU): plain probe, ncu launch list of one batch,
# one `ncu --set full` capture of bootbin_kernel (32 resamples = one group, N = 200k).
cd /root/repo
O=gpurun_out
TAG=${TAG:-r2m}
PFRAC=0.5,0.5,0.3,0.2,0.12,0.05 timeout 600 python tools/bootbin_probe.py > $O/bootbin_regimes_$TAG.log 2>&1 || echo "regime probe failed"
timeout 300 python tools/bootbin_probe.py > $O/bootbin_probe_$TAG.log 2>&1 || { echo "plain probe failed"; exit 1; }
tail -3 $O/bootbin_probe_$TAG.log
POLD=0 timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_bootbin_$TAG.csv \
    python tools/bootbin_probe.py > $O/ncu_launches_bootbin_$TAG.log 2>&1 || echo "launch list failed"
PN=200000 PB=32 POLD=0 timeout 400 ncu --set full --clock-control none --import-source on -k regex:bootbin_kernel -s 2 -c 1 \
    -o $O/prof_bootbin_$TAG -f python tools/bootbin_probe.py > $O/ncu_bootbin_$TAG.log 2>&1 || echo "ncu capture failed"
ls -la $O/prof_bootbin_$TAG.ncu-rep

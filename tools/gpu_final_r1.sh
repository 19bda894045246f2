# Final round-1 record: full check, ncu capture of the pair-by-pair (roofline) kernel, launch list of the 2PCF bench.
cd /root/repo
TAG=r1p bash tools/gpu_full_check.sh
PB_BLOCK_SUMS=0 PB_N=200000 PB_REPS=3 python tools/pb_run.py > gpurun_out/pp_plain_r1p.log 2>&1 && \
PB_BLOCK_SUMS=0 PB_N=200000 PB_REPS=2 ncu --set full --clock-control none --import-source on -k regex:pairbin_kernel -s 1 -c 1 \
  -o gpurun_out/prof_pairbin_pp_r1p -f python tools/pb_run.py > gpurun_out/ncu_pp_r1p.log 2>&1
tail -2 gpurun_out/pp_plain_r1p.log
python bench.py --skip-gp --skip-cpu --skip-other --npoints 200000 --steps 2 --warmup 3 > gpurun_out/bench_200k_r1p.json 2>/dev/null && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1p.csv \
  python bench.py --skip-gp --skip-cpu --skip-other --npoints 200000 --steps 2 --warmup 3 > gpurun_out/ncu_launch_r1p.log 2>&1
wc -l gpurun_out/launches_r1p.csv

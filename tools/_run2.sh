cd /root/repo
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_r2_cfg4_final.csv python bench.py --workload cfg4 --steps 3 --warmup 2 --no-cpu-baseline --no-cpu-parity > gpurun_out/r3l_ncu_l.log 2>&1; echo "launch list rc=$?"
ncu --set full --import-source on --clock-control none -k regex:hybrid_tile_fast -s 3 -c 1 -o gpurun_out/prof_r2_hyb8 -f python bench.py --workload cfg4 --steps 3 --warmup 1 --no-cpu-parity --no-cpu-baseline > gpurun_out/r3l_ncu_f.log 2>&1; echo "full rc=$?"
ncu --set full --clock-control none -k regex:finish_kernel -s 4 -c 1 -o gpurun_out/prof_r2_finish -f python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras --no-cpu-parity > gpurun_out/r3l_ncu_g.log 2>&1; echo "finish rc=$?"
ls -la gpurun_out/prof_r2_hyb8.ncu-rep gpurun_out/prof_r2_finish.ncu-rep gpurun_out/launches_r2_cfg4_final.csv | awk '{print $5,$9}'

cd /root/repo
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras --no-cpu-parity > gpurun_out/r3b_plain.json 2> gpurun_out/r3b_plain.err; echo "plain rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r2_cfg2.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras --no-cpu-parity > gpurun_out/r3b_ncu_l.log 2>&1; echo "launch list rc=$?"
ncu --set full --import-source on --clock-control none -k regex:scan_umma -s 4 -c 1 -o gpurun_out/prof_r2_umma -f python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras --no-cpu-parity > gpurun_out/r3b_ncu_f.log 2>&1; echo "full rc=$?"
ls -la gpurun_out/prof_r2_umma.ncu-rep gpurun_out/launches_r2_cfg2.csv

cd /root/repo
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8"
( time $T --steps 20 --warmup 5 > gpurun_out/r2u_bench_n8.json 2> gpurun_out/r2u_bench_n8.err ) 2> gpurun_out/r2u_time_n8.txt; echo "rc=$?"
$T --steps 2000 --warmup 50 --no-extras --no-cpu-parity --no-cpu-baseline > gpurun_out/r2u_bench_n8_long.json 2> /dev/null; echo "rc=$?"
RASS_B200_NO_OVERLAP=1 RASS_B200_SCAN_RESERVE=0 $T --steps 2000 --warmup 50 --no-extras --no-cpu-parity --no-cpu-baseline > gpurun_out/r2u_bench_n8_long_serial.json 2> /dev/null; echo "rc=$?"
RASS_B200_SCAN_RESERVE=4 $T --steps 2000 --warmup 50 --no-extras --no-cpu-parity --no-cpu-baseline > gpurun_out/r2u_bench_n8_long_res4.json 2> /dev/null; echo "rc=$?"
cat gpurun_out/r2u_time_n8.txt
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2u_bench_n8*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, round(d['value']), round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value']), 'kernel_ms', round(d['roofline']['kernel_ms'],4), 'sust', d.get('sustained',{}).get('qps'), d['clocks'], d['parity'].get('ids_equal_cpu_oracle'), (d.get('cfg5') or {}).get('qps'))
    except Exception as e: print(f, 'ERR', e)
PY
grep -c "NCCL INFO" gpurun_out/r2u_bench_n8.err; grep "nranks" gpurun_out/r2u_bench_n8.err | head -3

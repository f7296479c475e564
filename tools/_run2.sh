cd /root/repo
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 4"
$T --steps 20 --warmup 5 > gpurun_out/r3i_bench_n4.json 2> gpurun_out/r3i_bench_n4.err; echo "rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r3i_bench_n4.json').read().strip().splitlines()[-1])
print(round(d['value']), round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value']), d['roofline']['frac'], d['parity'].get('ids_equal_cpu_oracle'), d['clocks']['sm_mhz'])
print('cfg5', {k:d['cfg5'][k] for k in ('qps','ms_per_batch','scan_tflops_per_gpu','frac_of_tensor_peak')}, d['cfg5']['parity'].get('ids_equal_cpu_oracle'))
PY

cd /root/repo
python bench.py --workload cfg4 --steps 40 --warmup 5 > gpurun_out/r3g_cfg4.json 2> gpurun_out/r3g_cfg4.err; echo "rc=$?"; tail -2 gpurun_out/r3g_cfg4.err
python -c "
import json; d=json.loads(open('gpurun_out/r3g_cfg4.json').read().strip().splitlines()[-1]); print(round(d['value']), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), d['roofline']['kernels'], d['parity']['fused_ids_equal_cpu_oracle'], d['cpu_baseline']['value'], d['ingest']['tokens_per_s'])"

cd /root/repo
( time python bench.py --workload cfg3 --steps 6 --warmup 3 > gpurun_out/r3o_cfg3.json 2> gpurun_out/r3o_cfg3.err ) 2> gpurun_out/r3o_time.txt; echo "rc=$?"; tail -3 gpurun_out/r3o_time.txt
python -c "
import json; d=json.loads(open('gpurun_out/r3o_cfg3.json').read().strip().splitlines()[-1]); print(round(d['value']), round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value']), d['roofline'], d['parity'].get('ids_equal_cpu_oracle'), d['clocks'], d.get('sustained'))"

cd /root/repo
python -c "import __graft_entry__ as g; g.build(); g.smoke()" 2>&1 | tail -1
python -m pytest tests -m gpu -x -q > gpurun_out/r3n_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r3n_pytest.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r3n_bench.json 2> gpurun_out/r3n_bench.err; echo "bench rc=$?"
python -c "
import json; d=json.loads(open('gpurun_out/r3n_bench.json').read().strip().splitlines()[-1]); print(round(d['value']), round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value']), d['roofline']['frac'], d['parity']['ids_equal_cpu_oracle'], 'hybrid', round(d['hybrid']['qps']), d['hybrid']['parity']['fused_ids_equal_cpu_oracle'], 'b1', round(d['batch1']['qps']), {k:round(v['qps']) for k,v in d['batched'].items()})"

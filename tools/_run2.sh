cd /root/repo
python -m pytest tests -m gpu -x -q > gpurun_out/r3k_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r3k_pytest.log
python bench.py --workload cfg4 --steps 40 --warmup 5 > gpurun_out/r3k_cfg4.json 2> gpurun_out/r3k_cfg4.err; echo "rc=$?"
python -c "
import json; d=json.loads(open('gpurun_out/r3k_cfg4.json').read().strip().splitlines()[-1]); print(round(d['value']), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), 'text', round(d['roofline']['kernels']['bm25_fusion']['ms'],4), d['parity'])"

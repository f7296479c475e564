cd /root/repo
python -m pytest tests/test_gpu_text_ingest.py -x -q > gpurun_out/r2q_ingest.log 2>&1; echo "ingest rc=$?"; tail -30 gpurun_out/r2q_ingest.log
python -m pytest tests -m gpu -x -q > gpurun_out/r2q_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2q_pytest.log
python bench.py --workload cfg4 --steps 20 --warmup 5 > gpurun_out/r2q_cfg4.json 2> gpurun_out/r2q_cfg4.err; echo "cfg4 rc=$?"; tail -3 gpurun_out/r2q_cfg4.err
python -c "
import json; d=json.loads(open('gpurun_out/r2q_cfg4.json').read().strip().splitlines()[-1]); print(d['value'], d['ingest'], d['parity'], d['setup_s'])"

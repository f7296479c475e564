cd /root/repo
python -m pytest tests/test_gpu_lifecycle.py -x -q > gpurun_out/r2z_life.log 2>&1; echo "lifecycle rc=$?"; tail -15 gpurun_out/r2z_life.log
python -m pytest tests -m gpu -x -q > gpurun_out/r2z_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2z_pytest.log
python bench.py --steps 20 --warmup 5 --no-extras --no-cpu-baseline > gpurun_out/r2z_bench.json 2> gpurun_out/r2z_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/r2z_bench.err
python -c "
import json; d=json.loads(open('gpurun_out/r2z_bench.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['parity'].get('ids_equal_cpu_oracle'))"
nvidia-smi --query-gpu=memory.used --format=csv

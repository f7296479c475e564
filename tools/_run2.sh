cd /root/repo
python -m pytest tests/test_gpu_text_ingest.py -x -q > gpurun_out/r2x_ingest.log 2>&1; echo "ingest rc=$?"; tail -25 gpurun_out/r2x_ingest.log
python -m pytest tests/test_gpu_hybrid.py tests/test_gpu_hostquery.py tests/test_gpu_lifecycle.py -x -q 2>&1 | tail -3
python bench.py --workload cfg4 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2x_cfg4.json 2> gpurun_out/r2x_cfg4.err; echo "cfg4 rc=$?"; tail -2 gpurun_out/r2x_cfg4.err
python -c "
import json; d=json.loads(open('gpurun_out/r2x_cfg4.json').read().strip().splitlines()[-1]); print(d['value'], d['ingest'], d['parity']['fused_ids_equal_cpu_oracle'])"
B="python bench.py --no-extras --no-cpu-parity --no-cpu-baseline --steps 300 --warmup 20"
for rep in 1 2; do
$B > gpurun_out/rw_poll_$rep.json 2>/dev/null
RASS_DEBUG_RELAXED_WAIT=1 $B > gpurun_out/rw_relaxed_$rep.json 2>/dev/null
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/rw_*.json')):
    d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, round(d['value']), round(d['ms_per_step'],4), 'sust', round(d['sustained']['qps']), d['clocks']['sm_mhz'])
PY

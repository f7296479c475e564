cd /root/repo
python -m pytest tests/test_gpu_hybrid.py tests/test_gpu_text_ingest.py tests/test_gpu_hostquery.py tests/test_gpu_sharded_handle.py -x -q > gpurun_out/r2t_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2t_pytest.log
B="python bench.py --workload cfg4 --steps 40 --warmup 5 --no-cpu-baseline"
$B > gpurun_out/r2t_cfg4_limb_dyn.json 2> gpurun_out/r2t_cfg4.err; echo "rc=$?"; tail -2 gpurun_out/r2t_cfg4.err
RASS_HYBRID_NO_LIMB=1 $B --no-cpu-parity > gpurun_out/r2t_cfg4_nolimb_dyn.json 2>/dev/null
RASS_HYBRID_STATIC_CHUNKS=1 $B --no-cpu-parity > gpurun_out/r2t_cfg4_limb_static.json 2>/dev/null
RASS_HYBRID_NO_LIMB=1 RASS_HYBRID_STATIC_CHUNKS=1 $B --no-cpu-parity > gpurun_out/r2t_cfg4_nolimb_static.json 2>/dev/null
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2t_cfg4_*.json')):
    d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, round(d['value']), round(d['ms_per_step'],3), 'text', round(d['roofline']['kernels']['bm25_fusion']['ms'],4), 'text_only', round(d['text_only_kernel_ms'],4), d['parity'])
PY

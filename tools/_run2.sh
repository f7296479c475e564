cd /root/repo
python -m pytest tests -m gpu -x -q > gpurun_out/r2w_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2w_pytest.log
python tools/bench_client.py 300000 > gpurun_out/r2w_client.json 2> gpurun_out/r2w_client.err; echo "client rc=$?"; tail -2 gpurun_out/r2w_client.err; cat gpurun_out/r2w_client.json
( time python bench.py > gpurun_out/r2w_bench.json 2> gpurun_out/r2w_bench.err ) 2> gpurun_out/r2w_time.txt; echo "bench rc=$?"; cat gpurun_out/r2w_time.txt
( time python bench.py --impl reference > gpurun_out/r2w_ref.json 2> gpurun_out/r2w_ref.err ) 2>> gpurun_out/r2w_time.txt; echo "ref rc=$?"; tail -4 gpurun_out/r2w_time.txt
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2w_bench.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','e2e','gpu_launches','clocks')})
print(d['roofline']); print(d['parity']); print(d['sustained']); print(d['batch1']); print({k:(v if not isinstance(v,dict) else {kk:vv for kk,vv in v.items() if kk in ('qps','frac_of_tensor_peak','scan_tflops_per_gpu')}) for k,v in d['batched'].items()}); print({k:d['hybrid'][k] for k in ('qps','qps_e2e','ms_per_batch','knn_scan_ms','bm25_fusion_ms','frac_of_knn_ceiling','parity','ingest')}); print(d['cpu_baseline'])
r=json.loads(open('gpurun_out/r2w_ref.json').read().strip().splitlines()[-1]); print(r)
PY

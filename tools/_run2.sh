cd /root/repo
python -m pytest tests/test_gpu_hybrid.py tests/test_gpu_hostquery.py tests/test_gpu_lifecycle.py tests/test_gpu_text_ingest.py -x -q 2>&1 | tail -3
python tools/bench_client.py 300000 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(round(d['ingest_docs_per_s']), {k:round(v,3) for k,v in d.items() if k.endswith('_ms')})"

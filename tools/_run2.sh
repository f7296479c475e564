cd /root/repo
python -m pytest tests/test_gpu_sharded_handle.py tests/test_gpu_text_ingest.py tests/test_gpu_hybrid.py -x -q > gpurun_out/r3c_sharded.log 2>&1; echo "rc=$?"; tail -25 gpurun_out/r3c_sharded.log

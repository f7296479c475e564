cd /root/repo
python -m pytest tests/test_gpu_sharded.py tests/test_gpu_sharded_handle.py -x -q 2>&1 | tail -3

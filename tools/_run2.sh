cd /root/repo
for rows in 2500000 10000000; do
for ov in 1 0 1 0; do
python tools/bench_sharded.py --rows $rows --overlap $ov --steps 400 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print($rows, 'overlap', d['async_overlap'], round(d['qps']), round(d['ms_per_step'],4), d['ids_equal_cpu_oracle'], d['batches_repeated'])"
done
done

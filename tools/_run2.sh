cd /root/repo
for rep in 1 2 3; do
python bench.py --steps 20 --warmup 5 --no-extras --no-cpu-baseline --no-cpu-parity 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['value']), round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value']), round(d['e2e']['ms_per_step'],4), 'sust', round(d['sustained']['qps']), d['roofline']['frac'])"
done

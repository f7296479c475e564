cd /root/repo
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --no-extras --no-cpu-parity --no-cpu-baseline"
for rep in 1 2; do
for r in 0 1 2 4 8; do
RASS_B200_SCAN_RESERVE=$r $T --rows 2500000 --steps 1000 --warmup 20 > gpurun_out/n2r_2500000_res${r}_$rep.json 2>/dev/null
done
done
for r in 0 4; do
RASS_B200_SCAN_RESERVE=$r $T --rows 10000000 --steps 200 --warmup 20 > gpurun_out/n2r_10000000_res${r}_1.json 2>/dev/null
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/n2r_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, round(d['value']), round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value']), 'kernel_ms', round(d['roofline']['kernel_ms'],4), 'sust', round(d['sustained']['qps']), d['clocks']['sm_mhz'], d['clocks']['reasons'])
    except Exception as e: print(f, 'ERR', e)
PY

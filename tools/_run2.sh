cd /root/repo
B="python bench.py --no-extras --no-cpu-parity --no-cpu-baseline"
for rep in 1 2; do
for st in 6 5 4; do
RASS_DEBUG_UMMA_STAGES=$st $B --rows 10000000 --steps 100 --warmup 10 > gpurun_out/st_10M_${st}_$rep.json 2>/dev/null
RASS_DEBUG_UMMA_STAGES=$st $B --rows 1250000 --steps 800 --warmup 20 > gpurun_out/st_1M_${st}_$rep.json 2>/dev/null
done
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/st_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, round(d['value']), round(d['ms_per_step'],4), 'kernel_ms', round(d['roofline']['kernel_ms'],4), 'sust', round(d['sustained']['qps']), d['clocks']['sm_mhz'], d['parity']['fast_path_ids_equal_fp64_scan'])
    except Exception as e: print(f, 'ERR', e)
PY

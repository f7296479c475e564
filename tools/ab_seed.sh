# A/B of the threshold seed on one box: bash tools/ab_seed.sh [rows ...]
for rows in ${@:-10000000}; do
for i in 1 2; do
for m in seed noseed; do
  if [ $m = noseed ]; then export RASS_DEBUG_NO_SEED=1; else unset RASS_DEBUG_NO_SEED; fi
  python bench.py --rows $rows --steps 300 --warmup 20 --no-cpu-baseline --no-extras 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$m', d['config']['rows'], 'qps', round(d['value']), 'e2e', round(d['e2e']['value']), 'ms/step', round(d['ms_per_step'],4), 'scan ms', round(d['roofline']['kernel_ms'],4), 'MHz', d['clocks']['sm_mhz'], d['parity']['certificate_fallbacks_in_timed_region'])"
done; done; done

cd /root/repo
for r in list mask auto; do
if [ $r = auto ]; then unset RASS_B200_FILTER_ROUTE; else export RASS_B200_FILTER_ROUTE=$r; fi
python tools/bench_client.py 300000 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$r', {k:round(v,3) for k,v in d.items() if k.endswith('_ms')})"
done

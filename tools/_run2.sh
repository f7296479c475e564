cd /root/repo
cat > tools/_mb.py <<'PY'
import time, sys, os, numpy as np
sys.path.insert(0, os.getcwd())
import rassengine_b200 as rb
e = rb.Engine(dim=64, device=0)
e.append(np.random.default_rng(0).standard_normal((120000, 64)).astype(np.float32))
rng = np.random.default_rng(1)
def bulk(lo, hi, field, V):
    rows = np.arange(lo, hi)
    ip = np.arange(0, (hi - lo) * 24 + 1, 24, dtype=np.int64)
    t = rng.integers(0, V, size=(hi - lo) * 24).astype(np.int32)
    t0 = time.perf_counter(); e.text_add_rows(field, rows, ip, t); return time.perf_counter() - t0
ta = [bulk(0, 100000, f, 3000) for f in range(6)]
t0 = time.perf_counter(); e.text_commit([3000] * 6, 100000); tc = time.perf_counter() - t0
print("add 100k x6 fields ms", [round(x * 1e3, 1) for x in ta], "commit ms", round(tc * 1e3, 1), flush=True)
tb = [bulk(100000, 110000, f, 3000) for f in range(6)]
t0 = time.perf_counter(); e.text_commit([3000] * 6, 110000); tc2 = time.perf_counter() - t0
print("add 10k x6 fields ms", [round(x * 1e3, 1) for x in tb], "commit ms", round(tc2 * 1e3, 1), flush=True)
e.close()
PY
RASS_DEBUG_TEXT_TIMES=1 python tools/_mb.py 2>&1

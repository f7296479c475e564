"""Quick device-side probe: fill N synthetic rows on the GPU, time each scan path.  python tools/probe.py [N] [paths]"""
import sys
import time

import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import rassengine_b200 as rb

N = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
which = sys.argv[2].split(",") if len(sys.argv) > 2 else ["stream", "umma"]
NOEXACT = len(sys.argv) > 3
D = 1024
e = rb.Engine(dim=D, capacity_rows=N)
g = torch.Generator(device="cuda").manual_seed(1234)
t0 = time.time()
CH = 500_000
for c0 in range(0, N, CH):
    m = min(CH, N - c0)
    x = torch.randn((m, D), generator=g, device="cuda", dtype=torch.float32)
    x /= x.norm(dim=1, keepdim=True) + 1e-9
    torch.cuda.synchronize()
    e.append_dev(x.data_ptr(), m)
print(f"fill {N} rows: {time.time() - t0:.1f}s", flush=True)
gq = torch.Generator(device="cuda").manual_seed(5678)
CASES = (("stream", 1, 10), ("stream", 2, 10), ("umma", 64, 10), ("umma", 1, 10), ("umma", 64, 100), ("stream", 1, 100),
         ("gemm", 256, 10), ("gemm", 1024, 10), ("gemm", 8192, 100), ("gemm", 200, 10))
custom = [w.split(":") for w in which if ":" in w]      # e.g. gemm:1024:10
CLUSTERED = set()
if custom:
    CASES = tuple((c[0], int(c[1]), int(c[2])) for c in custom)     # a 4th field "c" = clustered queries
    CLUSTERED = {(c[0], int(c[1]), int(c[2])) for c in custom if len(c) > 3 and c[3] == "c"}
    which = [c[0] for c in custom]
for path, B, k in CASES:
    if path not in which:
        continue
    e.set_path(getattr(rb, "PATH_" + path.upper()))
    q = torch.randn((B, D), generator=gq, device="cuda")
    if (path, B, k) in CLUSTERED:
        # queries = stored rows + N(0, 0.05^2) noise per element (SURVEY.md 8d "clustered variant"): near-ties at the top
        pick = torch.randint(0, N, (B,), generator=gq, device="cuda")
        base = torch.from_numpy(e.read_rows(0, 1)).cuda()   # placeholder to keep dtype/device
        rows_np = np.stack([e.read_rows(int(r), 1)[0] for r in pick.tolist()])
        q = torch.from_numpy(rows_np).cuda() + 0.05 * torch.randn((B, D), generator=gq, device="cuda") / 32.0
    rows = torch.empty((B, k), dtype=torch.int64, device="cuda")
    sc = torch.empty((B, k), dtype=torch.float32, device="cuda")
    for it in range(int(os.environ.get("PROBE_ITERS", "5"))):
        st = e.search_knn_dev(q.data_ptr(), B, k, rows.data_ptr(), sc.data_ptr())
    gbs = st["bytes_streamed"] / (st["scan_ms"] * 1e-3) / 1e9 if st["scan_ms"] else 0
    print(path, "B", B, "k", k, {kk: (round(v, 3) if isinstance(v, float) else v) for kk, v in st.items()},
          f"scan {gbs:.0f} GB/s  qps {B / (st['total_ms'] * 1e-3):.1f}", flush=True)
    if path == "gemm":
        # every query against the 64-per-pass tcgen05 path (both certified exact, so ids must be identical)
        e.set_path(rb.PATH_UMMA)
        rows3 = torch.empty((B, k), dtype=torch.int64, device="cuda")
        st3 = e.search_knn_dev(q.data_ptr(), B, k, rows3.data_ptr(), sc.data_ptr())
        print("   ids equal to the 64-per-pass path:", bool((rows == rows3).all()), "umma total ms", round(st3["total_ms"], 2),
              "fallbacks", st3["n_fallback"], f"tensor {2.0 * B * N * D / (st['scan_ms'] * 1e-3) / 1e12:.0f} TFLOP/s", flush=True)
    if NOEXACT:
        continue
    # parity of the fast path against the fp64 scan on the device (ids must be identical)
    e.set_path(rb.PATH_EXACT)
    nb = min(B, 4)
    rows2 = torch.empty((nb, k), dtype=torch.int64, device="cuda")
    sc2 = torch.empty((nb, k), dtype=torch.float32, device="cuda")
    st2 = e.search_knn_dev(q.data_ptr(), nb, k, rows2.data_ptr(), sc2.data_ptr())
    print("   exact-scan ids equal:", bool((rows[:nb] == rows2).all()), "exact ms", round(st2["total_ms"], 2), flush=True)

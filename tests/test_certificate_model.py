"""A numpy model of the scan -> pool -> certificate scheme, checked for the one property the build
rests on: WHENEVER the certificate holds, the re-ranked candidates are the exact top-k -- for any data, any segment
assignment, any pivot the compactions or the threshold seed choose, clusters of near-duplicates included.  The model
follows the kernels' rules, not their code: rows are dealt to segments in tiles, a segment appends keys above its
threshold, compacts to `keep..2 keep` entries above a pivot once it passes SEG / 2, picks up the largest pivot any
segment has published between tiles, and starts from the sampled seed (store.cu).  finish: t = max threshold used,
b' = k-th best approximate key in the pool, candidates >= b' - 2 eps re-ranked exactly, certified iff t + eps < s_k."""
import numpy as np
from hypothesis import given, settings, strategies as st

from oracle import knn

NEG = -np.inf


def _bf16(a):
    u = np.ascontiguousarray(a, dtype=np.float32).view(np.uint32).astype(np.uint64)
    u = (u + 0x7FFF + ((u >> 16) & 1)) & 0xFFFF0000
    return u.astype(np.uint32).view(np.float32).reshape(np.shape(a))


def _scan_keys(X, q):
    """approximate keys of the bf16 scan and the allowance eps of finish.cu (cosine)."""
    X64, q64 = X.astype(np.float64), q.astype(np.float64)
    xn = np.sqrt((X64 * X64).sum(axis=1))
    q_hat = (q64 / np.sqrt((q64 * q64).sum())).astype(np.float32)
    X16, q16 = _bf16(X), _bf16(q_hat)
    rho_x = float((np.sqrt(((X64 - X16.astype(np.float64)) ** 2).sum(axis=1)) / xn).max())
    qh = q_hat.astype(np.float64)
    rho_q = float(np.sqrt(((qh - q16.astype(np.float64)) ** 2).sum()) / np.sqrt((qh * qh).sum()))
    eps = np.float32((rho_x * (1.0 + rho_q) + rho_q + X.shape[1] * 2.0 ** -23) * 1.0001)
    key = (X16 @ q16).astype(np.float32) * (1.0 / xn).astype(np.float32)
    return key.astype(np.float32), float(eps)


def _model_search(key, exact, k, n_segs, tile, seg, keep, seed_rows, seed_rank, rng):
    n = key.size
    thr = np.full(n_segs, NEG, dtype=np.float32)
    pools = [[] for _ in range(n_segs)]                  # (key, row)
    gthr = NEG
    if seed_rows is not None and seed_rows.size >= seed_rank:
        s = np.sort(key[seed_rows])[::-1]
        v = s[seed_rank - 1]                             # rank-th largest sampled key
        gthr = np.nextafter(np.float32(v), np.float32(NEG))    # strictly below it: `rank` keys lie strictly above
        thr[:] = gthr
    n_tiles = (n + tile - 1) // tile
    for t in range(n_tiles):
        sgm = t % n_segs
        g_seen = gthr                                    # read before the tile, applied after it (as in the kernels)
        rows = np.arange(t * tile, min(n, (t + 1) * tile))
        hit = rows[key[rows] > thr[sgm]]
        pools[sgm].extend((key[r], int(r)) for r in hit)
        thr[sgm] = max(thr[sgm], g_seen)
        if len(pools[sgm]) > seg // 2:
            ks = np.array([p[0] for p in pools[sgm]], dtype=np.float32)
            order = np.sort(ks)[::-1]
            # any pivot with keep..2*keep entries strictly above it; ties may leave fewer (the upper end is used)
            want = int(rng.integers(keep, 2 * keep + 1))
            pivot = order[min(want, order.size - 1)]
            pools[sgm] = [p for p in pools[sgm] if p[0] > pivot]
            thr[sgm] = max(thr[sgm], pivot)
            gthr = max(gthr, pivot)
    return pools, thr


def _finish(pools, thr, key_eps, exact, k):
    eps = key_eps
    entries = [p for pl in pools for p in pl]
    t = float(thr.max())
    if not entries:
        return None, False
    ks = np.array([p[0] for p in entries], dtype=np.float32)
    rows = np.array([p[1] for p in entries], dtype=np.int64)
    if ks.size >= k:
        bprime = np.sort(ks)[::-1][k - 1]
        cand = rows[ks >= bprime - 2 * eps]
    else:
        cand = rows
    ex = exact[cand]
    order = np.lexsort((cand, -ex))[:k]
    top = cand[order]
    if top.size < k:
        return top, t == NEG                              # fewer than k candidates: exact only if nothing was excluded
    s_k = ex[order][k - 1]
    return top, (t == NEG) or (t + eps < s_k)


@settings(max_examples=120, deadline=None, derandomize=True)
@given(st.integers(0, 2**31 - 1), st.sampled_from([1, 10, 32, 100]), st.sampled_from([2, 5, 16]),
       st.sampled_from([(256, 32), (512, 128)]), st.booleans(), st.sampled_from(["iid", "cluster", "dups"]))
def test_certified_means_exact(seed, k, n_segs, seg_keep, seeded, kind):
    rng = np.random.default_rng(seed)
    n, d = int(rng.integers(300, 4000)), 32
    X = rng.standard_normal((n, d)).astype(np.float32)
    q = rng.standard_normal(d).astype(np.float32)
    if kind == "cluster":                                # adjacent near-duplicates of the query's neighbourhood
        lo = int(rng.integers(0, n - 200))
        X[lo:lo + 200] = (q[None, :] + 0.05 * rng.standard_normal((200, d))).astype(np.float32)
    elif kind == "dups":                                 # exact duplicates: ties on the key, rows decide
        X[rng.integers(0, n, size=40)] = X[int(rng.integers(0, n))]
    key, eps = _scan_keys(X, q)
    exact = knn.cos64(X, q)
    seg, keep = seg_keep
    seed_rows = rng.choice(n, size=min(n, 256), replace=False) if seeded else None
    pools, thr = _model_search(key, exact, k, n_segs, 64, seg, keep, seed_rows, keep + 8, rng)
    top, certified = _finish(pools, thr, eps, exact, k)
    want = np.lexsort((np.arange(n), -exact))[:k]
    if certified:
        assert top.tolist() == want.tolist()
    # the scheme is not vacuous: on iid data with room above k it certifies
    if kind == "iid" and keep >= k and n >= 2000:
        assert certified

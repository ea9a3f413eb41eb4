import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..", "tests"))
import numpy as np
import flye_b200 as fb

def ed_np(a, b):
    n, m = len(a), len(b)
    prev = np.arange(m + 1, dtype=np.int32)
    for i in range(1, n + 1):
        sub = prev[:-1] + (b != a[i - 1])
        dele = prev[1:] + 1
        cur = np.minimum(sub, dele)
        # insertion needs a running min: cur[j] = min(cur[j], cur[j-1] + 1)
        cur = np.concatenate([[i], cur]).astype(np.int32)
        idx = np.arange(m + 1, dtype=np.int32)
        cur = np.minimum.accumulate(cur - idx) + idx
        prev = cur
    return int(prev[m])

eng = fb.Engine(0)
rng = np.random.default_rng(3)
bad = 0
for t in range(2):
    n = int(rng.integers(2500, 6000))
    a = rng.integers(0, 4, n).astype(np.uint8)
    err = [0.03, 0.06, 0.045, 0.02][t % 4]
    b = []
    for c in a:
        u = rng.random()
        if u < err / 3: continue
        if u < 2 * err / 3: b.append(rng.integers(0, 4)); b.append(c); continue
        b.append((c + 1) % 4 if u < err else c)
    b = np.array(b, dtype=np.uint8)
    d = ed_np(a, b)
    row = [n, len(b), d]
    for limit in (None,):
        if t == 0: continue
        if limit is None: os.environ.pop("FG_DEBUG_ED_LIMIT", None)
        else: os.environ["FG_DEBUG_ED_LIMIT"] = str(limit)
        got = eng.debug_edit_distance(a, b)
        ok = (got == d) if (limit is None or d < limit) else (got >= limit)
        row.append((limit, got, "ok" if ok else "BAD"))
        bad += not ok
    print(row, flush=True)
print("bad =", bad)

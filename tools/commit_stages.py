"""Stage timers of PolynomialBatch.from_values at the City proof shape (2^12 x 135, pinned / pageable).  Development tool."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import city_rollup_b200 as m
from util import rand_felts
c = m.Context(0)
for log_n, n_cols in ((12, 135), (12, 20), (13, 135)):
    n = 1 << log_n
    pinned = c.pinned_empty((n_cols, n))
    pinned[:] = rand_felts(5, (n_cols, n))
    pageable = np.array(pinned)
    for name, src in (("pinned", pinned), ("pageable", pageable)):
        for rep in range(3):
            b = m.PolynomialBatch.from_values(c, src, 3, False, 4); b.cap; b.free()
        c.profile_enable(True); c.profile_read()
        t0 = time.perf_counter()
        for rep in range(10):
            b = m.PolynomialBatch.from_values(c, src, 3, False, 4); b.cap; b.free()
        wall = (time.perf_counter() - t0) / 10 * 1e3
        st, cnt = c.profile_read(); c.profile_enable(False)
        print(log_n, n_cols, name, "wall ms %.3f" % wall, {k: round(v / 10, 3) for k, v in st.items() if v})

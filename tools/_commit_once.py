"""development helper: a few from_values commits of device-resident random columns (for ncu launch lists / captures)
usage: _commit_once.py [log_rows=20] [n_cols=135] [reps=2]"""
import sys

sys.path[:0] = ['.', 'tests', 'tools']
import torch

import city_rollup_b200 as m

log_n = int(sys.argv[1]) if len(sys.argv) > 1 else 20
n_cols = int(sys.argv[2]) if len(sys.argv) > 2 else 135
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
c = m.Context(0)
g = torch.Generator(device="cuda")
g.manual_seed(7)
t = torch.empty((n_cols, 1 << log_n), dtype=torch.int64, device="cuda")
t.random_(0, 2**62, generator=g)
torch.cuda.synchronize()
for i in range(reps):
    c.profile_enable(True)
    c.profile_read()
    c.timer_start()
    b = m.PolynomialBatch.from_values_device(c, t.data_ptr(), n_cols, log_n, 3, 4)
    ms = c.timer_stop_ms()
    st, cnt = c.profile_read()
    c.profile_enable(False)
    b.free()
    print("commit 2^%d x %d: %.3f ms  stages %s" % (log_n, n_cols, ms, {k: round(v, 3) for k, v in st.items() if v}), file=sys.stderr)

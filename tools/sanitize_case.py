#!/usr/bin/env python
"""tools/sanitize_case.py -- a small invocation of every kernel, for compute-sanitizer (memcheck / racecheck):
    compute-sanitizer --tool memcheck python tools/sanitize_case.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

import channelcoding_b200 as cc  # noqa: E402

ctx = cc.Context(0)
rng = np.random.default_rng(0)
for q, t, rows in ((4, 2, None), (6, 5, None), (6, 5, 63), (7, 10, None), (8, 18, None)):
    code = ctx.bch(q, errors=t)
    if rows:
        code.set_rows(rows)
    y = (1 + 0.7 * rng.standard_normal((257, code.n))).astype(np.float32)
    for variant in ("NMS", "SCMS2", "2DNMS", "SPA"):
        code.decode(y, variant, 0.8, 0.9, 8, stop_rule=1)
    code.awgn_point(4.0, 1000, "NMS", 0.8)
    code.bitflip_point(2, "MS", count=300)
    code.awgn_point_hard(5.0, 500)
c = ctx.bch(6, dmin=7)
g = ctx.from_dense(c.H()[rng.permutation(c.h_rows)], c.rate)
g.decode((1 + 0.7 * rng.standard_normal((64, 63))).astype(np.float32), "OMS", 1.0, 0.01, 10)
rs = ctx.rs(8, 16)
w = rs.encode(rng.integers(0, 256, size=(300, rs.l)).astype(np.uint8))
w[:, 5] ^= 7
w[::3, 100:110] ^= 1
rs.gf_decode(w)
ep = np.zeros((300, 4), np.uint8)
ep[:, 0] = 5
rs.gf_decode(w, erasures=(ep, np.ones(300, np.uint8)))
ctx.awgn_llr(63, 0.7, 0, 0, 0, 1001)
ctx.close()
print("sanitize case done")

#!/usr/bin/env python
"""tools/profile_targets.py -- launches each secondary kernel a few times (for ncu captures and quick
timings): K1 awgn_llr, K4 gf_decode RS(255,223), K2c ms_cyclic_cta BCH(255,131), K2g ms_csr on H_alt."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def timed(fn, reps=3):
    import torch
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


def main():
    import numpy as np
    import torch
    import channelcoding_b200 as cc
    ctx = cc.Context(0)
    ctx.use_torch_stream()
    # K1: 4M frames x 63 floats = 1.06 GB written
    frames = 1 << 22
    y = torch.empty((frames, 63), dtype=torch.float32, device="cuda")
    ms = timed(lambda: ctx.awgn_llr(63, 0.7, 0, 1, 0, frames, out=y))
    print("K1 awgn_llr n=63: %.3f ms, %.1f GB/s written (%.3e symbols/s)" % (ms, frames * 63 * 4 / ms / 1e6, frames * 63 / ms * 1e3))
    # K4
    rs = ctx.rs(8, 16)
    rng = np.random.default_rng(5)
    words = rs.encode(rng.integers(0, 256, size=(4096, rs.l)).astype(np.uint8))
    bad = words.copy()
    for i in range(4096):
        ne = i % 18
        pos = rng.choice(255, ne, replace=False)
        bad[i, pos] ^= rng.integers(1, 256, size=ne).astype(np.uint8)
    count = 1 << 20
    d = torch.from_numpy(bad).cuda().repeat(count // 4096, 1).contiguous()
    out = (torch.empty_like(d), torch.empty(count, dtype=torch.uint8, device="cuda"), torch.empty(count, dtype=torch.uint8, device="cuda"))
    ms = timed(lambda: rs.gf_decode(d, out=out))
    print("K4 gf_decode RS(255,223), 0..17 errors: %.3f ms, %.3e words/s, %.1f GB/s algorithmic" % (ms, count / ms * 1e3, count * 511 / ms / 1e6))
    clean = torch.from_numpy(words).cuda().repeat(count // 4096, 1).contiguous()
    ms = timed(lambda: rs.gf_decode(clean, out=out))
    print("K4 gf_decode RS(255,223), error free: %.3f ms, %.3e words/s, %.1f GB/s algorithmic" % (ms, count / ms * 1e3, count * 511 / ms / 1e6))
    # K2c
    c255 = ctx.bch(8, errors=18)
    cnt = torch.zeros(8, dtype=torch.int64, device="cuda")
    fr = 200000
    ms = timed(lambda: c255.awgn_point(6.0, fr, "NMS", 0.8, out=cnt))
    print("K2c ms_cyclic_cta BCH(255,131) NMS 6 dB: %.3f ms, %.3e frames/s" % (ms, fr / ms * 1e3))
    # K2g on the reference's H_alt shape: here the (63,45) parity checks in a non-cyclic row order
    c = ctx.bch(6, dmin=7)
    H = c.H()[np.random.default_rng(1).permutation(c.h_rows)]
    g = ctx.from_dense(H, c.rate)
    fr = 200000
    ms = timed(lambda: g.awgn_point(4.0, fr, "NMS", 0.8, out=cnt))
    print("K2g ms_csr (63,45) permuted H NMS 4 dB: %.3f ms, %.3e frames/s" % (ms, fr / ms * 1e3))
    ms = timed(lambda: c.awgn_point(4.0, fr * 10, "NMS", 0.8, out=cnt))
    print("K2  ms_cyclic (63,45) NMS 4 dB: %.3f ms, %.3e frames/s" % (ms, fr * 10 / ms * 1e3))
    c15 = ctx.bch(4, errors=2)
    ms = timed(lambda: c15.awgn_point(3.0, 2000000, "SPA", max_iter=50, stop_rule=1, out=cnt))
    print("K2g ms_csr BCH(15,7) SPA 3 dB (y not scaled): %.3f ms, %.3e frames/s" % (ms, 2000000 / ms * 1e3))
    ctx.close()


if __name__ == "__main__":
    main()

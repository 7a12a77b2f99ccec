#!/usr/bin/env python
"""tools/ab_fused.py -- fused Monte-Carlo point rate (ccgpu_awgn_point) of one code over a list of Eb/N0 values, frame
count calibrated to about 60 ms per launch.  CCGPU_LIB selects the library under test (A/B runs)."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--q", type=int, default=6)
    ap.add_argument("--t", type=int, default=5)
    ap.add_argument("--ebno", type=float, nargs="+", default=[4.0, 6.0, 7.0, 8.0, 10.0])
    ap.add_argument("--variant", default="NMS")
    ap.add_argument("--alpha", type=float, default=0.8)
    ap.add_argument("--rows", type=int, default=0)
    ap.add_argument("--quick", type=int, default=-1)
    ap.add_argument("--lane", type=int, default=-1)
    ap.add_argument("--reps", type=int, default=5)
    a = ap.parse_args()
    import torch
    import channelcoding_b200 as cc
    ctx = cc.Context(0)
    ctx.use_torch_stream()
    ctx.set_option("quick", a.quick)
    ctx.set_option("lane", a.lane)
    code = ctx.bch(a.q, errors=a.t)
    if a.rows:
        code.set_rows(a.rows)
    lib = os.environ.get("CCGPU_LIB", "libccgpu.so").split("/")[-1]
    for eb in a.ebno:
        cnt = torch.zeros(8, dtype=torch.int64, device="cuda")

        def run(frames):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            cnt.zero_()
            e0.record()
            code.awgn_point(eb, frames, a.variant, a.alpha, 0.0, 50, out=cnt, seed=1, point=int(eb * 2))
            e1.record()
            torch.cuda.synchronize()
            return e0.elapsed_time(e1)
        frames = 1 << 18
        run(frames)
        ms = run(frames)
        frames = int(min(max(frames * 60.0 / max(ms, 1e-3), 1 << 18), 1 << 31))
        best = min(run(frames) for _ in range(a.reps))
        c = cnt.cpu().numpy()
        print("%s quick=%d lane=%d (%d,%d) rows=%d %s %5.1f dB: %.4e frames/s  (%d frames, %.2f ms)  wer %.3e it %.4f  counters %s" % (
            lib, a.quick, a.lane, code.n, code.l, code.h_rows, a.variant, eb, frames / best * 1e3, frames, best, c[1] / c[0], c[3] / c[0],
            " ".join(str(int(x)) for x in c[:6])), flush=True)


if __name__ == "__main__":
    main()

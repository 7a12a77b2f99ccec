#!/usr/bin/env python
"""tools/ab_ms.py -- quick kernel timing for A/B experiments: resident-LLR decode and fused point of one
code (default BCH(63,36) NMS at 4 dB).  CCGPU_LIB selects the library under test."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--q", type=int, default=6)
    ap.add_argument("--t", type=int, default=5)
    ap.add_argument("--ebno", type=float, default=4.0)
    ap.add_argument("--frames", type=int, default=1 << 21)
    ap.add_argument("--variant", default="NMS")
    ap.add_argument("--alpha", type=float, default=0.8)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--rows", type=int, default=0, help="redundant H: this many cyclic shifts (ccgpu_code_set_rows)")
    ap.add_argument("--stop", type=int, default=0)
    a = ap.parse_args()
    import numpy as np
    import torch
    import channelcoding_b200 as cc
    ctx = cc.Context(0)
    ctx.use_torch_stream()
    code = ctx.bch(a.q, errors=a.t)
    if a.rows:
        code.set_rows(a.rows)
    y = torch.empty((a.frames, code.n), dtype=torch.float32, device="cuda")
    ctx.awgn_llr(code.n, np.float32(cc.sigma(code.rate, a.ebno)), 0, 1, 0, a.frames, out=y)
    out = (torch.empty((a.frames, code.n), dtype=torch.uint8, device="cuda"), None,
           torch.empty(a.frames, dtype=torch.uint8, device="cuda"), torch.empty(a.frames, dtype=torch.uint8, device="cuda"))
    best = 1e9
    for _ in range(a.reps + 2):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        code.decode(y, a.variant, a.alpha, 0.0, 50, a.stop, out=out, want_L=False)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    cnt = torch.zeros(8, dtype=torch.int64, device="cuda")
    bestf = 1e9
    for _ in range(a.reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        code.awgn_point(a.ebno, a.frames, a.variant, a.alpha, 0.0, 50, a.stop, out=cnt)
        e1.record()
        torch.cuda.synchronize()
        bestf = min(bestf, e0.elapsed_time(e1))
    it = torch.where(out[3] == 1, torch.full_like(out[2], 50).int(), out[2].int() + 1).double().mean().item()
    print("%s n=%d rows=%d %s %.1f dB: %.2f it, %.3e edge-it/s resident | resident %.4e frames/s (%.3f ms)   fused %.4e frames/s   checksum %d" % (
        os.environ.get("CCGPU_LIB", "libccgpu.so").split("/")[-1], code.n, code.h_rows, a.variant, a.ebno, it,
        a.frames / best * 1e3 * it * code.edges, a.frames / best * 1e3, best,
        a.frames / bestf * 1e3, int(out[2].sum().item()) + int(out[3].sum().item())))


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""tools/bench_fixed.py -- fixed-point (two frames per lane) vs float32 min-sum: fused Monte-Carlo points on the same
Philox noise, frames/s and edge-iterations/s side by side.  One JSON line per (code, Eb/N0).

    python tools/bench_fixed.py [--frames N] [--quick]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--reps", type=int, default=3)
    a = ap.parse_args()
    import torch
    import channelcoding_b200 as cc
    ctx = cc.Context(0)
    ctx.use_torch_stream()

    def timed(code, eb, frames, variant, alpha, quant):
        out = torch.zeros(8, dtype=torch.int64, device="cuda")
        code.awgn_point(eb, max(1000, frames // 20), variant, alpha, 0.0, 50, out=out, quant=quant)
        best = 1e30
        for _ in range(a.reps):
            out.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            code.awgn_point(eb, frames, variant, alpha, 0.0, 50, seed=1, point=7, out=out, quant=quant)
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        c = out.cpu().numpy()
        return {"frames_per_s": frames / best * 1e3, "ms": best, "wer": float(c[1]) / frames,
                "avg_iterations": float(c[3]) / frames,
                "edge_iterations_per_s": float(c[3]) / best * 1e3 * code.edges}

    cases = [("BCH(15,7)", (4, 2), "MS", 1.0, (8.0, 31, 31), (1.0, 3.0, 6.0), 4e7),
             ("BCH(31,16)", (5, 3), "NMS", 0.8, (8.0, 31, 31), (3.0, 6.0), 2e7),
             ("BCH(63,36)", (6, 5), "NMS", 0.8, (8.0, 31, 31), (2.0, 4.0, 6.0, 8.0), 1.6e7),
             ("BCH(63,36) q=(16,63,63)", (6, 5), "NMS", 0.8, (16.0, 63, 63), (4.0,), 1.6e7),
             ("BCH(127,64)", (7, 10), "NMS", 0.8, (8.0, 31, 31), (3.0, 5.0, 7.0), 4e6),
             ("BCH(255,131)", (8, 18), "NMS", 0.8, (8.0, 31, 29), (4.0, 6.0, 8.0), 4e5)]
    for name, (q, t), variant, alpha, quant, ebnos, frames in cases:
        code = ctx.bch(q, errors=t)
        frames = int(frames * (0.1 if a.quick else 1.0))
        for eb in ebnos:
            f = timed(code, eb, frames, variant, alpha, None)
            x = timed(code, eb, frames, variant + "_Q", alpha, quant)
            print(json.dumps({"config": name, "ebno_db": eb, "variant": variant, "quant": quant, "frames": frames,
                              "float": f, "fixed": x, "speedup_frames": x["frames_per_s"] / f["frames_per_s"],
                              "speedup_edge_iterations": x["edge_iterations_per_s"] / f["edge_iterations_per_s"]}), flush=True)
    ctx.close()


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""tools/group_bench.py -- multi-GPU behind the C ABI (ccgpu_group): per-point overhead and STRONG scaling.

One process, no torch.distributed: every Eb/N0 point is one ccgpu_group_awgn_point call that shards the global frame
range over the member devices (one host thread per device) and returns the merged counters.  Prints JSON lines:

  latency   a point of `frames` frames (1 .. 1e6) on groups of 1/2/4/8 devices: wall time per call = launch + wait +
            counter merge; the part that does not shrink with more devices
  sweep     the reference's own schedule (simulation.c++:91-93,105-112: N = min(1e6, 5e3 / WER_prev) frames per point,
            0.5 dB steps from the Shannon-limit start to 8 dB), total wall time on 1/2/4/8 devices (strong scaling)
  point     one large point (1e8 frames) on 1/2/4/8 devices
  waterfall BASELINE config 5: BCH(255,131) 0 dB upwards until the curve is below 1e-7

    python tools/group_bench.py --gpus 1,2,4,8 [--modes latency,sweep,point,waterfall]
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", default="1")
    ap.add_argument("--modes", default="latency,sweep,point")
    ap.add_argument("--variant", default="NMS")
    ap.add_argument("--waterfall-variant", default="NMS")
    ap.add_argument("--waterfall-to", type=float, default=12.0)
    ap.add_argument("--waterfall-target", type=float, default=1e-7)
    ap.add_argument("--waterfall-max-frames", type=float, default=2e10)
    ap.add_argument("--min-errors", type=int, default=100)
    a = ap.parse_args()
    import channelcoding_b200 as cc
    L = cc._lib.lib()
    sizes = [int(x) for x in a.gpus.split(",")]
    modes = a.modes.split(",")
    quant = (8.0, 31, 31)

    for n in sizes:
        g = cc.Group(n)
        code = g.bch(6, errors=5)
        variants = [("NMS", None), ("NMS_Q", quant)]
        code.awgn_point(4.0, 100000 * n, "NMS", 0.8)  # warm-up: module load, first launch on every device
        code.awgn_point(4.0, 100000 * n, "NMS_Q", 0.8, quant=quant)
        if "latency" in modes:
            for frames in (1, 1000, 16384 * n, 100000, 1000000):
                for minf in ((16384, 1) if frames < 16384 * n else (16384,)):
                    g.set_min_frames(minf)
                    best, tot = 1e9, 0.0
                    reps = 30
                    for _ in range(reps):
                        t0 = time.perf_counter()
                        c = code.awgn_point(4.0, frames, "NMS", 0.8, seed=1, point=3)
                        dt = time.perf_counter() - t0
                        best, tot = min(best, dt), tot + dt
                    print(json.dumps({"mode": "latency", "n_gpus": n, "frames": frames, "min_frames_per_member": minf,
                                      "best_us": best * 1e6, "mean_us": tot / reps * 1e6, "frames_counted": c["frames"]}), flush=True)
            g.set_min_frames(16384)
        if "sweep" in modes:
            for variant, q in variants:
                start = L.ccgpu_sweep_start_ebno(code.rate, 0.5)
                for cap in (10 ** 6, 10 ** 8):
                    t0 = time.perf_counter()
                    wer, point, eb, frames_total, pts = 0.5, 0, start, 0, []
                    while eb < max(8.0, start) + 0.25:
                        nfr = int(L.ccgpu_sweep_samples(wer, cap))
                        c = code.awgn_point(eb, nfr, variant, 0.8, seed=0, point=point, quant=q)
                        wer = c["frame_errors"] / nfr
                        pts.append((eb, nfr, c["frame_errors"]))
                        if wer == 0.0:
                            wer = 5e3 / cap
                        frames_total += nfr
                        point += 1
                        eb += 0.5
                    el = time.perf_counter() - t0
                    print(json.dumps({"mode": "sweep", "n_gpus": n, "variant": variant, "cap": cap, "points": len(pts),
                                      "frames": frames_total, "seconds": el, "frames_per_s": frames_total / el,
                                      "checksum": sum(p[2] for p in pts)}), flush=True)
        if "point" in modes:
            for variant, q in variants:
                for eb in (4.0,):
                    frames = 10 ** 8
                    t0 = time.perf_counter()
                    c = code.awgn_point(eb, frames, variant, 0.8, seed=2, point=1, quant=q)
                    el = time.perf_counter() - t0
                    print(json.dumps({"mode": "point", "n_gpus": n, "variant": variant, "ebno_db": eb, "frames": frames,
                                      "seconds": el, "frames_per_s": frames / el, "wer": c["frame_errors"] / frames}), flush=True)
        if "waterfall" in modes:
            big = g.bch(8, errors=18)
            wq = (8.0, 31, 29) if a.waterfall_variant.endswith("_Q") else None
            eb, point = 0.0, 0
            while eb <= a.waterfall_to + 1e-9:
                tot = {"frames": 0, "frame_errors": 0, "bit_errors": 0, "iterations": 0, "failures": 0, "undetected": 0}
                batch = 1 << 20
                t0 = time.perf_counter()
                while True:
                    c = big.awgn_point(eb, batch, a.waterfall_variant, 0.8, seed=0, point=point, frame0=tot["frames"], quant=wq)
                    for k in tot:
                        tot[k] += c[k]
                    if tot["frame_errors"] >= a.min_errors or tot["frames"] >= a.waterfall_max_frames:
                        break
                    batch = min(batch * 4, 1 << 30)
                el = time.perf_counter() - t0
                wer = tot["frame_errors"] / tot["frames"]
                print(json.dumps({"mode": "waterfall", "code": big.to_string(a.waterfall_variant), "n_gpus": n, "ebno_db": eb,
                                  "frames": tot["frames"], "frame_errors": tot["frame_errors"], "wer": wer,
                                  "ber": tot["bit_errors"] / tot["frames"] / big.n,
                                  "avg_iterations": tot["iterations"] / tot["frames"], "seconds": el,
                                  "frames_per_s": tot["frames"] / el}), flush=True)
                if wer < a.waterfall_target and tot["frame_errors"] >= 10:
                    break
                if tot["frame_errors"] == 0:
                    break
                point += 1
                eb += 0.5
        g.close()


if __name__ == "__main__":
    main()

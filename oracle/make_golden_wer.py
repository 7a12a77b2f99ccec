#!/usr/bin/env python
"""oracle/make_golden_wer.py -- TEST INFRASTRUCTURE.  Word-error statistics of the REFERENCE itself for the statistical
leg of the parity tests: oracle/_ref/libccref.so (the reference compiled by build_ref.sh) runs the inner loop of
awgn_simulation::operator() (simulation.c++:124-136: mt19937_64 + normal_distribution -> decoder.correct -> word-error
test) for >= 1e5 frames per point, three points per code.  The counts go to tests/golden/ref_wer.json; the GPU tests
compare the engine's Philox-driven WER with them (binomial interval on the reference's finite sample).

    python oracle/make_golden_wer.py [--frames 100000]      (about 15 min on 8 cores; the (127,64) points dominate)
"""
import argparse
import json
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)

# (name, q, errors, variant id of ref_shim.cc, variant, alpha, points)
CASES = [("bch_15_7", 4, 2, 0, "MS", 1.0, (1.0, 3.0, 5.0)),
         ("bch_63_36", 6, 5, 1, "NMS", 0.8, (2.0, 4.0, 5.0)),
         ("bch_127_64", 7, 10, 1, "NMS", 0.8, (4.0, 5.0, 6.0))]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=100000)
    a = ap.parse_args()
    import ccref
    ref = ccref.Ref()
    threads = os.cpu_count() or 1
    per = (a.frames + threads - 1) // threads
    out = {"source": "oracle/_ref/libccref.so = the reference (REF-FIXED) running simulation.c++:124-136; seeds 1000..1000+T-1",
           "threads": threads, "points": []}
    for name, q, t, vid, variant, alpha, points in CASES:
        for eb in points:
            t0 = time.time()
            frames, werr, el = ref.awgn_baseline(ccref.FAM_BCH, q, ccref.CAP_ERRORS, t, ccref.ALG_SOFT0 + vid, eb, seed=1000,
                                                 seconds=1e9, threads=threads, max_frames_per_thread=per)
            out["points"].append({"code": name, "q": q, "errors": t, "variant": variant, "alpha": alpha, "max_iter": 50,
                                  "ebno_db": eb, "frames": frames, "word_errors": werr})
            print(name, eb, frames, werr, werr / frames, "%.0f s" % (time.time() - t0), flush=True)
            with open(os.path.join(HERE, "..", "tests", "golden", "ref_wer.json"), "w") as f:
                json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""tools/comparison_h.py -- the experiment behind the reference's comparison_h.pdf: word error
rate of the six min-sum variants on BCH(63,45,7) and BCH(127,106,7) decoded on

  * "H"          the k cyclic shifts of the reversed check polynomial   (codes/cyclic.h:346-359), kernel K2
  * "H_alt"      "H from the roots of g(x)", exponents reduced mod n     (the construction as intended), kernel K2g
  * "H_alt_ref"  the same with the reference's from_power (mod 2^q, codes/cyclic.h:361-385 + galois.h:182-184,
                 SURVEY defect C4) -- not a parity-check matrix of the code for exponents >= 2^q

with the reference's variant constants (src/benchmark.c++ catalogue).  One JSON line per (code, matrix, variant,
Eb/N0); the points run on the fused Monte-Carlo path (Philox channel + decode + counters on the device).

    python tools/comparison_h.py [--frames 2000000] [--ebno-from 2 --ebno-to 8 --ebno-step 1]
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

# name -> (variant, alpha, beta): the default constants of the reference's tags at HEAD (soft_decision.h:16-60),
# i.e. what `min_sum(H, y, nms_tag{})` etc. run with (oracle/ccref.py VARIANT_PARAMS ids 0-5)
VARIANTS = {
    "MS": ("MS", 1.0, 0.0),
    "NMS": ("NMS", 0.8, 0.0),
    "OMS": ("OMS", 1.0, 0.01),
    "SCMS1": ("SCMS1", 1.0, 0.0),
    "SCMS2": ("SCMS2", 1.0, 0.0),
    "2DNMS": ("2DNMS", 1.0, 1.0),
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=2_000_000)
    ap.add_argument("--min-errors", type=int, default=200)
    ap.add_argument("--ebno-from", type=float, default=2.0)
    ap.add_argument("--ebno-to", type=float, default=8.0)
    ap.add_argument("--ebno-step", type=float, default=1.0)
    ap.add_argument("--variants", default=",".join(VARIANTS))
    ap.add_argument("--seed", type=int, default=0)
    a = ap.parse_args()
    import channelcoding_b200 as cc
    ctx = cc.Context(0)
    for q, dmin in ((6, 7), (7, 7)):
        base = ctx.bch(q, dmin=dmin)
        mats = {"H": base,
                "H_alt": ctx.from_dense(base.H_alt(as_reference=False), base.rate),
                "H_alt_ref": ctx.from_dense(base.H_alt(as_reference=True), base.rate)}
        name = base.to_string("")[:-1]
        for mname, code in mats.items():
            for vname in a.variants.split(","):
                variant, alpha, beta = VARIANTS[vname]
                eb, point = a.ebno_from, 0
                while eb < a.ebno_to + a.ebno_step / 2:
                    t0 = time.perf_counter()
                    done, errs, bits, iters = 0, 0, 0, 0
                    # rounds of growing size until enough word errors or the frame budget is spent
                    batch = max(10000, a.frames // 16)
                    while done < a.frames and errs < a.min_errors:
                        nf = min(batch, a.frames - done)
                        c = code.awgn_point(eb, nf, variant, alpha, beta, 50, 0, seed=a.seed, point=point, frame0=done)
                        done += c["frames"]
                        errs += c["frame_errors"]
                        bits += c["bit_errors"]
                        iters += c["iterations"]
                        batch *= 2
                    print(json.dumps({"code": name, "matrix": mname, "rows": int(code.H().shape[0]), "variant": vname,
                                      "alpha": alpha, "beta": beta, "ebno_db": eb, "frames": done, "word_errors": errs,
                                      "wer": errs / done, "ber": bits / done / code.n, "avg_iterations": iters / done,
                                      "kernel": {1: "ms_cyclic", 2: "ms_csr"}[code.kernel],
                                      "seconds": time.perf_counter() - t0}), flush=True)
                    eb += a.ebno_step
                    point += 1


if __name__ == "__main__":
    main()

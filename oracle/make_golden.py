#!/usr/bin/env python
"""oracle/make_golden.py -- TEST INFRASTRUCTURE.  Regenerates tests/golden/*.

Runs the REFERENCE ITSELF (oracle/_ref/libccref.so = hannesweisbach/channelcoding compiled by
oracle/build_ref.sh, REF-FIXED flavour) on seeded inputs and stores input/output pairs as small
fixtures.  Only runnable where /root/reference was available to build oracle/_ref; the fixtures
are committed so the GPU box and CI never need the reference.

  python oracle/make_golden.py            # rewrites tests/golden/
"""
import itertools
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ccref  # noqa: E402
from ccref import (ALG_BM, ALG_EUKLID, ALG_PGZ, ALG_SOFT0, CAP_DMIN, CAP_ERRORS, FAM_BCH, FAM_RS,  # noqa: E402
                   VARIANT_PARAMS)

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")

# (name, family, q, cap kind, cap value, t)
CODES = [
    ("bch_15_7", FAM_BCH, 4, CAP_ERRORS, 2, 2),
    ("bch_15_5", FAM_BCH, 4, CAP_DMIN, 7, 3),
    ("bch_15_7_dmin5", FAM_BCH, 4, CAP_DMIN, 5, 2),
    ("bch_15_7_dmin6", FAM_BCH, 4, CAP_DMIN, 6, 2),
    ("bch_31_26", FAM_BCH, 5, CAP_DMIN, 3, 1),
    ("bch_31_21", FAM_BCH, 5, CAP_DMIN, 5, 2),
    ("bch_31_16", FAM_BCH, 5, CAP_DMIN, 7, 3),
    ("bch_31_11", FAM_BCH, 5, CAP_DMIN, 9, 4),
    ("bch_63_57", FAM_BCH, 6, CAP_DMIN, 3, 1),
    ("bch_63_51", FAM_BCH, 6, CAP_DMIN, 5, 2),
    ("bch_63_45", FAM_BCH, 6, CAP_DMIN, 7, 3),
    ("bch_63_39", FAM_BCH, 6, CAP_DMIN, 9, 4),
    ("bch_63_36", FAM_BCH, 6, CAP_ERRORS, 5, 5),
    ("bch_127_120", FAM_BCH, 7, CAP_DMIN, 3, 1),
    ("bch_127_113", FAM_BCH, 7, CAP_DMIN, 5, 2),
    ("bch_127_106", FAM_BCH, 7, CAP_DMIN, 7, 3),
    ("bch_127_99", FAM_BCH, 7, CAP_DMIN, 9, 4),
    ("bch_127_64", FAM_BCH, 7, CAP_ERRORS, 10, 10),
    ("bch_255_131", FAM_BCH, 8, CAP_ERRORS, 18, 18),
    ("rs_7_5", FAM_RS, 3, CAP_ERRORS, 1, 1),
    ("rs_7_3", FAM_RS, 3, CAP_ERRORS, 2, 2),
    ("rs_15_9", FAM_RS, 4, CAP_ERRORS, 3, 3),
    ("rs_255_223", FAM_RS, 8, CAP_ERRORS, 16, 16),
]


def sigma(rate, ebno_db):
    # simulation.c++:83-85
    return float(np.float32(1.0) / np.sqrt(2 * rate * 10 ** (ebno_db / 10.0)))


def special_frames(y, rng):
    """rows that exercise exact ties / zeros / erasure-like inputs."""
    n = y.shape[1]
    y[0, :] = 1.0                      # clean all-zero word
    y[1, :] = 1.0; y[1, n // 3] = -1.0  # one hard flip, all magnitudes tied
    y[2, :] = 0.0                      # all erased
    y[3, :] = 1.0; y[3, 1] = 0.0; y[3, n // 2] = 0.0  # two zeros
    y[4, :] = 1.0; y[4, 2] = 0.0       # a single zero
    y[5, :] = -1.0                     # all-one word (a codeword of every primitive BCH code here? no: decodes oddly)
    y[6, :] = 1.0; y[6, rng.choice(n, 2, replace=False)] = -1.0
    y[7, :] = np.where(rng.random(n) < 0.5, 0.5, -0.5).astype(np.float32)
    return y


def main():
    os.makedirs(OUT, exist_ok=True)
    r = ccref.Ref()
    assert r.flavour == "REF-FIXED"
    rng = np.random.default_rng(20261018)

    # ---- code catalogue + GF tables ------------------------------------------------------
    cat = {}
    arrays = {}
    for name, fam, q, kind, val, t in CODES:
        p = r.params(fam, q, kind, val)
        entry = dict(family=fam, q=q, cap_kind=kind, cap_value=val, **p)
        assert p["t"] == t
        entry["to_string"] = {a: r.to_string(fam, q, kind, val, aid)
                              for a, aid in (("PGZ", ALG_PGZ), ("BM", ALG_BM), ("EUKLID", ALG_EUKLID))}
        cat[name] = entry
        arrays[name + ".g"] = r.poly(fam, q, kind, val, "g").astype(np.uint16)
        arrays[name + ".h"] = r.poly(fam, q, kind, val, "h").astype(np.uint16)
        if fam == FAM_BCH:
            arrays[name + ".H"] = np.packbits(r.H(fam, q, kind, val), axis=1)
            arrays[name + ".H_alt"] = np.packbits(r.H(fam, q, kind, val, alt=True), axis=1)
    # soft to_string spot checks
    for name, fam, q, kind, val in (("bch_15_7", FAM_BCH, 4, CAP_ERRORS, 2), ("bch_31_16", FAM_BCH, 5, CAP_DMIN, 7),
                                    ("bch_63_36", FAM_BCH, 6, CAP_ERRORS, 5)):
        for v in range(6):
            cat[name]["to_string"][VARIANT_PARAMS[v][0]] = r.to_string(fam, q, kind, val, ALG_SOFT0 + v)
    for q in range(1, 9):
        e, l = r.gf_tables(q)
        arrays["gf%d.exp" % q] = e
        arrays["gf%d.log" % q] = l
    np.savez_compressed(os.path.join(OUT, "codes.npz"), **arrays)

    # ---- min-sum fixtures: min_sum<float,uint8_t>(H, y, Tag{}) ---------------------------------
    ms_cases = [("bch_15_7", (1.0, 3.0, 6.0), 192), ("bch_31_16", (2.0, 5.0), 128),
                ("bch_63_36", (2.0, 4.0, 6.0), 96), ("bch_127_64", (3.0, 5.0), 24), ("bch_255_131", (5.0,), 10)]
    by_name = {c[0]: c for c in CODES}
    for name, ebnos, frames in ms_cases:
        _, fam, q, kind, val, t = by_name[name]
        p = cat[name]
        H = r.H(fam, q, kind, val)
        ys = []
        for eb in ebnos:
            y = (1.0 + sigma(p["rate"], eb) * rng.standard_normal((frames, p["n"]))).astype(np.float32)
            ys.append(y)
        y = np.concatenate(ys)
        y = special_frames(y, rng)
        out = {"y": y, "ebno": np.repeat(np.asarray(ebnos, np.float32), frames)}
        variants = range(12) if p["n"] <= 127 else (0, 1, 4)
        for v in variants:
            bits, L, it, failed = r.min_sum(v, H, y)
            out["v%d.bits" % v] = np.packbits(bits, axis=1)
            out["v%d.iter" % v] = it.astype(np.uint8)
            out["v%d.failed" % v] = failed
            if v < 9:
                out["v%d.L" % v] = L
        np.savez_compressed(os.path.join(OUT, "minsum_%s.npz" % name), **out)
        # decoder.correct() path (cyclic.h:254-267) must agree with min_sum on H()
        if name in ("bch_15_7", "bch_31_16", "bch_63_36"):
            for v in range(6):
                b2, f2 = r.soft_correct(fam, q, kind, val, v, y)
                b1 = np.unpackbits(out["v%d.bits" % v], axis=1)[:, :p["n"]]
                assert np.array_equal(f2, out["v%d.failed" % v]) and np.array_equal(b1[f2 == 0], b2[f2 == 0])

    # ---- algebraic fixtures: code.correct<uint8_t>(word) with euklid_tag -----------------------
    hd_cases = [("rs_255_223", 160), ("rs_15_9", 400), ("rs_7_3", 300), ("rs_7_5", 200), ("bch_15_7", 400),
                ("bch_31_16", 400), ("bch_63_36", 300), ("bch_127_64", 120), ("bch_255_131", 60)]
    for name, count in hd_cases:
        _, fam, q, kind, val, t = by_name[name]
        p = cat[name]
        msgs = rng.integers(0, (1 << q) if fam == FAM_RS else 2, size=(count, p["l"])).astype(np.uint8)
        msgs[0, :] = 0
        msgs[1, :] = (1 << q) - 1 if fam == FAM_RS else 1
        words = r.encode(fam, q, kind, val, msgs)
        bad = words.copy()
        nerr = rng.integers(0, t + 4, size=count)
        nerr[:3] = (0, 0, t)
        for i in range(count):
            pos = rng.choice(p["n"], nerr[i], replace=False)
            if fam == FAM_RS:
                bad[i, pos] ^= rng.integers(1, 1 << q, size=nerr[i]).astype(np.uint8)
            else:
                bad[i, pos] ^= 1
        fixed, status = r.hard_correct(fam, q, kind, val, ALG_EUKLID, bad)
        np.savez_compressed(os.path.join(OUT, "hard_%s.npz" % name), msgs=msgs, words=words, received=bad,
                            nerr=nerr.astype(np.uint8), corrected=fixed, status=status)

    # ---- known answers -----------------------------------------------------------------------
    kat = {}
    # Table 3 / bitflips.c++: exhaustive +-1 patterns on (31,16,7); failures per weight
    fam, q, kind, val = FAM_BCH, 5, CAP_DMIN, 7
    H = r.H(fam, q, kind, val)
    table = {}
    for w in range(0, 4):
        pats = list(itertools.combinations(range(31), w))
        y = np.ones((len(pats), 31), np.float32)
        for i, pp in enumerate(pats):
            y[i, list(pp)] = -1.0
        row = {"patterns": len(pats)}
        for v in range(6):
            bits, L, it, failed = r.min_sum(v, H, y)
            row[VARIANT_PARAMS[v][0]] = int(((failed == 1) | bits.any(axis=1)).sum())
        words = (y < 0).astype(np.uint8)
        for a, aid in (("BM", ALG_BM), ("PGZ", ALG_PGZ), ("EUKLID", ALG_EUKLID)):
            fixed, status = r.hard_correct(fam, q, kind, val, aid, words)
            row[a] = int(((status != 0) | fixed.any(axis=1)).sum())
        table[str(w)] = row
    kat["bitflip_31_16_7"] = table
    # report p.33 Table 3, percent failures for w = 0..6 (published; MS/SCMS1/SCMS2/BM are parameter free)
    kat["table3_percent"] = {"MS": [0, 0, 29.7, 79.2, 96.3, 99.5, 99.7], "SCMS1": [0, 0, 29.2, 76.1, 96.7, 99.8, 99.8],
                             "SCMS2": [0, 0, 2.4, 55.1, 95.3, 99.9, 100], "BM": [0, 0, 0, 0, 100, 100, 100]}

    # exercises.c++ tasks (vectors are the reference's own; results from running the reference)
    def run(fam, q, kind, val, alg, word, erasures=()):
        fixed, status = r.hard_correct(fam, q, kind, val, alg, np.asarray([word], np.uint8), erasures)
        return {"received": list(map(int, word)), "erasures": list(map(int, erasures)), "status": int(status[0]),
                "corrected": list(map(int, fixed[0]))}

    def rs_word(q, powers):  # Element::from_power(p) / Element(0) encoded as -1
        e, _ = r.gf_tables(q)
        return [0 if p < 0 else int(e[p]) for p in powers]

    ex = {}
    a61 = [1, 1, 1, 0, 0, 0, 1, 0, 0, 1, 1, 0, 1, 0, 1]
    ex["6.1.b1"] = dict(run(FAM_BCH, 4, CAP_DMIN, 7, ALG_PGZ, [1, 1, 1, 1, 0, 0, 1, 0, 0, 1, 0, 0, 1, 1, 1]), expect=a61,
                        code="bch_15_5", alg="PGZ")
    ex["6.1.b2"] = dict(run(FAM_BCH, 4, CAP_DMIN, 7, ALG_PGZ, [1, 1, 1, 1, 0, 0, 1, 0, 0, 1, 0, 0, 1, 0, 1]), expect=a61,
                        code="bch_15_5", alg="PGZ")
    ex["6.2"] = dict(run(FAM_BCH, 4, CAP_DMIN, 5, ALG_PGZ, [1, 0, 0, 1, 0, 1, 1, 1, 1, 0, 1, 1, 0, 0, 0]), expect=None,
                     code="bch_15_7_dmin5", alg="PGZ")
    a63 = [1, 1, 1, 1, 0, 1, 1, 1, 0, 1, 0, 0, 0, 1, 1]
    ex["6.3.b1"] = dict(run(FAM_BCH, 4, CAP_DMIN, 6, ALG_PGZ, [1, 1, 1, 1, 0, 1, 1, 1, 0, 1, 0, 0, 0, 0, 1]), expect=a63,
                        code="bch_15_7_dmin6", alg="PGZ")
    ex["6.3.b2"] = dict(run(FAM_BCH, 4, CAP_DMIN, 6, ALG_PGZ, [1, 1, 1, 1, 0, 1, 1, 1, 0, 1, 0, 0, 1, 0, 1]), expect=a63,
                        code="bch_15_7_dmin6", alg="PGZ")
    ex["6.3.b3"] = dict(run(FAM_BCH, 4, CAP_DMIN, 6, ALG_PGZ, [0, 0, 0, 1, 0, 1, 1, 1, 0, 1, 0, 0, 0, 1, 1]), expect="unspecified",
                        code="bch_15_7_dmin6", alg="PGZ")  # exercises.c++:99-105 accepts either outcome
    ex["6.4.b1"] = dict(run(FAM_RS, 3, CAP_ERRORS, 1, ALG_PGZ, rs_word(3, [-1, -1, -1, -1, -1, -1, 4])), expect=[0] * 7,
                        code="rs_7_5", alg="PGZ")
    ex["6.4.b2"] = dict(run(FAM_RS, 3, CAP_ERRORS, 1, ALG_PGZ, rs_word(3, [2, 2, 0, -1, -1, -1, 4])), expect="unspecified",
                        code="rs_7_5", alg="PGZ")
    ex["6.5"] = dict(run(FAM_RS, 4, CAP_ERRORS, 3, ALG_PGZ, [1, 1, 1, 1] + [0] * 11), expect="unspecified", code="rs_15_9",
                     alg="PGZ")
    ex["6.6"] = dict(run(FAM_RS, 3, CAP_ERRORS, 2, ALG_PGZ, rs_word(3, [6, 2, 2, 5, -1, -1, 5])),
                     expect=rs_word(3, [6, 2, 2, 5, 4, 6, 5]), code="rs_7_3", alg="PGZ")
    a67 = rs_word(3, [6, 2, 2, 5, 4, 6, 5])
    b67 = list(a67)
    for e_ in (5, 4, 3, 2):
        b67[e_] = 0
    ex["6.7"] = dict(run(FAM_RS, 3, CAP_ERRORS, 2, ALG_BM, b67, (5, 4, 3, 2)), expect=a67, code="rs_7_3", alg="BM")
    ex["6.8"] = dict(run(FAM_RS, 3, CAP_ERRORS, 2, ALG_BM, rs_word(3, [2, 0, 4, 0, 5, 0, 2]), (1, 3)),
                     expect=rs_word(3, [2, 5, 4, 6, 5, 6, 2]), code="rs_7_3", alg="BM")
    for a, aid in (("PGZ", ALG_PGZ), ("BM", ALG_BM)):
        ex["6.9." + a] = dict(run(FAM_RS, 3, CAP_ERRORS, 2, aid, rs_word(3, [3, 4, 0, 3, 4, 3, 3])),
                              expect=rs_word(3, [2, 4, 0, 3, 4, 3, 2]), code="rs_7_3", alg=a)
    for a, aid in (("PGZ", ALG_PGZ), ("BM", ALG_BM), ("EUKLID", ALG_EUKLID)):
        ex["6.10." + a] = dict(run(FAM_BCH, 4, CAP_ERRORS, 2, aid, [1, 0, 1, 0, 0, 1, 1, 1, 1, 0, 1, 1, 1, 1, 1]),
                               expect=[1, 0, 1, 0, 0, 1, 1, 1, 1, 0, 1, 0, 1, 0, 1], code="bch_15_7", alg=a)
    for k_, v_ in ex.items():  # the reference must reproduce its own published answers
        if v_["expect"] == "unspecified":
            continue
        if v_["expect"] is None:
            assert v_["status"] != 0, k_
        else:
            assert v_["status"] == 0 and v_["corrected"] == v_["expect"], (k_, v_)
    kat["exercises"] = ex
    with open(os.path.join(OUT, "catalogue.json"), "w") as f:
        json.dump(cat, f, indent=1, sort_keys=True)
    with open(os.path.join(OUT, "kat.json"), "w") as f:
        json.dump(kat, f, indent=1, sort_keys=True)
    total = sum(os.path.getsize(os.path.join(OUT, f)) for f in os.listdir(OUT))
    print("golden fixtures written to", OUT, "total bytes", total)


if __name__ == "__main__":
    main()

"""The C restatement (oracle/) against the committed golden vectors, which were produced by
running the reference itself (oracle/make_golden.py -> oracle/_ref/libccref.so, REF-FIXED).
CPU only; this is what pins the oracle on machines without /root/reference."""
import itertools

import numpy as np
import pytest

import oracle
from conftest import VARIANT_PARAMS, golden_H, load_golden

BCH = ["bch_15_7", "bch_15_5", "bch_15_7_dmin5", "bch_15_7_dmin6", "bch_31_26", "bch_31_21", "bch_31_16",
       "bch_31_11", "bch_63_57", "bch_63_51", "bch_63_45", "bch_63_39", "bch_63_36", "bch_127_120",
       "bch_127_113", "bch_127_106", "bch_127_99", "bch_127_64", "bch_255_131"]
RS = ["rs_7_5", "rs_7_3", "rs_15_9", "rs_255_223"]


@pytest.mark.parametrize("q", range(1, 9))
def test_gf_tables(q, golden_codes):
    exp, log = oracle.gf_tables(q)
    assert np.array_equal(exp, golden_codes["gf%d.exp" % q])
    assert np.array_equal(log, golden_codes["gf%d.log" % q])


@pytest.mark.parametrize("name", BCH + RS)
def test_code_construction(name, catalogue, golden_codes):
    e = catalogue[name]
    c = oracle.Code(e["family"], e["q"], e["t"])
    assert (c.n, c.l, c.k, c.dmin) == (e["n"], e["l"], e["k"], e["dmin"])
    assert c.rate == e["rate"]
    assert np.array_equal(c.poly("g"), golden_codes[name + ".g"])
    assert np.array_equal(c.poly("h")[:c.deg_h + 1], golden_codes[name + ".h"])
    if e["family"] == 0:
        assert np.array_equal(c.H(), golden_H(golden_codes, catalogue, name))


@pytest.mark.parametrize("name", ["bch_15_7", "bch_31_16", "bch_63_36", "bch_127_64", "bch_255_131"])
def test_min_sum_golden(name, catalogue, golden_codes):
    g = load_golden("minsum_%s.npz" % name)
    H = golden_H(golden_codes, catalogue, name)
    y = g["y"]
    n = catalogue[name]["n"]
    for v, (variant, alpha, beta, max_iter) in VARIANT_PARAMS.items():
        if "v%d.iter" % v not in g:
            continue
        sel = slice(None) if n <= 63 else slice(0, 12)  # the dense restatement is O(k w n) per iteration
        bits, L, it, failed = oracle.min_sum(H, y[sel], variant, alpha, beta, max_iter)
        assert np.array_equal(failed, g["v%d.failed" % v][sel]), (name, v)
        assert np.array_equal(it, g["v%d.iter" % v][sel]), (name, v)
        ok = failed == 0
        gb = np.unpackbits(g["v%d.bits" % v], axis=1)[sel, :n]
        assert np.array_equal(bits[ok], gb[ok]), (name, v)
        if "v%d.L" % v in g:
            assert np.array_equal(L[ok].view(np.uint32), g["v%d.L" % v][sel][ok].view(np.uint32)), (name, v)


@pytest.mark.parametrize("name", ["rs_255_223", "rs_15_9", "rs_7_3", "rs_7_5", "bch_15_7", "bch_31_16", "bch_63_36",
                                  "bch_127_64", "bch_255_131"])
def test_hard_golden(name, catalogue):
    g = load_golden("hard_%s.npz" % name)
    e = catalogue[name]
    c = oracle.Code(e["family"], e["q"], e["t"])
    assert np.array_equal(c.encode(g["msgs"]), g["words"])
    out, nerr, status = c.hard_correct(g["received"])
    assert np.array_equal(status, g["status"])
    ok = status == 0
    assert np.array_equal(out[ok], g["corrected"][ok])
    # bounded-distance property of the golden data itself
    within = g["nerr"] <= e["t"]
    assert ok[within].all() and np.array_equal(out[within], g["words"][within])
    assert ((out[ok] != g["received"][ok]).sum(axis=1) == nerr[ok]).all()


def test_bitflip_table3(kat, catalogue, golden_codes):
    """bitflips.c++ / report Table 3 on BCH(31,16,7): failures per error weight."""
    H = golden_H(golden_codes, catalogue, "bch_31_16")
    c = oracle.Code(0, 5, 3)
    for w in range(0, 4):
        pats = list(itertools.combinations(range(31), w))
        y = np.ones((len(pats), 31), np.float32)
        for i, pp in enumerate(pats):
            y[i, list(pp)] = -1.0
        row = kat["bitflip_31_16_7"][str(w)]
        assert row["patterns"] == len(pats)
        for v in (0, 1, 2, 3, 4, 5):
            variant, alpha, beta, max_iter = VARIANT_PARAMS[v]
            if w == 3 and v not in (0, 4):
                continue  # keep the CPU suite short
            bits, L, it, failed = oracle.min_sum(H, y, variant, alpha, beta, max_iter)
            assert int(((failed == 1) | bits.any(axis=1)).sum()) == row[variant], (w, variant)
        out, nerr, status = c.hard_correct((y < 0).astype(np.uint8))
        assert int(((status != 0) | out.any(axis=1)).sum()) == row["EUKLID"]
    # published percentages (report p.33) for the parameter-free decoders
    for variant in ("MS", "SCMS2"):
        for w in (2, 3):
            row = kat["bitflip_31_16_7"][str(w)]
            assert abs(100.0 * row[variant] / row["patterns"] - kat["table3_percent"][variant][w]) <= 0.1


def test_exercises(kat, catalogue):
    """exercises.c++ tasks 6.1-6.10 (errors-only and erasure cases) through the Euklid restatement."""
    for key, ex in kat["exercises"].items():
        e = catalogue[ex["code"]]
        c = oracle.Code(e["family"], e["q"], e["t"])
        if ex["alg"] == "PGZ" and ex["erasures"]:
            continue
        out, nerr, status = c.hard_correct(np.asarray([ex["received"]], np.uint8), ex["erasures"])
        if ex["expect"] == "unspecified":
            # bounded-distance decoders agree whenever both succeed
            if status[0] == 0 and ex["status"] == 0:
                assert list(map(int, out[0])) == ex["corrected"], key
            continue
        if ex["expect"] is None:
            assert status[0] != 0, key
        else:
            assert status[0] == 0 and list(map(int, out[0])) == ex["expect"], key


def test_philox_known_answers():
    """Random123 kat_vectors for philox4x32-10."""
    assert [int(x) for x in oracle.philox4x32_10([0] * 4, [0] * 2)] == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert [int(x) for x in oracle.philox4x32_10([0xffffffff] * 4, [0xffffffff] * 2)] == \
        [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert [int(x) for x in oracle.philox4x32_10([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344],
                                                 [0xa4093822, 0x299f31d0])] == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_channel_statistics():
    y = oracle.awgn(seed=0, point=3, frame0=0, frames=4000, n=63, sigma_f=0.7)
    z = (y.astype(np.float64) - 1.0) / 0.7
    assert abs(z.mean()) < 0.01 and abs(z.std() - 1.0) < 0.01
    assert abs((np.abs(z) > 1.959964).mean() - 0.05) < 0.003
    assert abs(oracle.sigma(36 / 63, 4.0) - 1 / np.sqrt(2 * 36 / 63 * 10 ** 0.4)) < 1e-7


def pgz_fill_restatement(code, t, received, epos, ecnt):
    """codes/bch.h:97-149 on top of the errors-only decoder: zero fill, one fill, fewer corrected positions wins
    (ties: zero fill), more than 2t erasures or two failures = decoding failure"""
    count, n = received.shape
    out = np.zeros_like(received)
    status = np.ones(count, np.uint8)
    fills = np.repeat(received, 2, axis=0)
    for i in range(count):
        fills[2 * i, epos[i, :ecnt[i]]] = 0
        fills[2 * i + 1, epos[i, :ecnt[i]]] = 1
    cand, nerr, st = code.hard_correct(fills)
    for i in range(count):
        if ecnt[i] > 2 * t:
            continue
        ok0, ok1 = st[2 * i] == 0, st[2 * i + 1] == 0
        if not (ok0 or ok1):
            continue
        pick = 0 if ok0 and (not ok1 or nerr[2 * i] <= nerr[2 * i + 1]) else 1
        out[i], status[i] = cand[2 * i + pick], 0
    return out, status


def test_pgz_erasure_rule_pinned_by_the_reference(catalogue):
    """vectors produced by the reference's own primitive_bch<.., peterson_gorenstein_zierler_tag>::correct(b, erasures)
    (oracle/make_golden_pgz.py) == the restatement above driven by the oracle's bounded-distance decoder"""
    g = load_golden("hard_pgz_erasures.npz")
    for name in ("bch_15_7", "bch_31_16", "bch_63_36", "bch_63_45"):
        e = catalogue[name]
        c = oracle.Code(e["family"], e["q"], e["t"])
        out, status = pgz_fill_restatement(c, e["t"], g[name + ".received"], g[name + ".epos"], g[name + ".ecnt"])
        # words on which the reference's PGZ linear solver fails inside the correction radius (its own TODO,
        # hard_decision.h:69-71; found for t = 5 only) say nothing about the erasure rule: left out, but counted
        clean = g[name + ".solver_defect"] == 0
        assert clean.mean() > 0.97, name
        assert np.array_equal(status[clean], g[name + ".status"][clean]), name
        ok = clean & (status == 0)
        assert np.array_equal(out[ok], g[name + ".corrected"][ok]), name


# ---- fixed-point restatement (unpinned extension): quantiser known answers and convergence to the pinned float decoder
def test_fixed_point_quantiser_kat():
    import ctypes as C
    q = oracle.lib().oracle_quantise
    q.restype = C.c_int
    assert q(0.0625, 8.0, 31) == 0 and q(0.1875, 8.0, 31) == 2 and q(-0.1875, 8.0, 31) == -2  # ties to even
    assert q(0.3125, 8.0, 31) == 2 and q(0.4375, 8.0, 31) == 4
    assert q(100.0, 8.0, 31) == 31 and q(-100.0, 8.0, 31) == -31
    assert q(float("inf"), 8.0, 31) == 31 and q(float("-inf"), 8.0, 31) == -31 and q(float("nan"), 8.0, 31) == 0
    assert q(1.0, 8.0, 31) == 8 and q(-1.0, 3.0, 7) == -3


def test_fixed_point_converges_to_float(catalogue):
    """with a fine quantiser and no saturation the integer loop takes the float decoder's hard
    decisions.  Stated disagreement: the decoder is chaotic on these dense matrices (a last-bit difference moves the
    iteration at which the all-zero word appears), so the bar is on the DECISIONS of frames that converge in both
    runs and on the word error rate, not on iteration indices."""
    rng = np.random.default_rng(0)
    for name, frames, eb in (("bch_15_7", 4000, 3.0), ("bch_63_36", 1500, 4.0)):
        e = catalogue[name]
        H = oracle.Code(0, e["q"], e["t"]).H()
        y = (1 + oracle.sigma(e["rate"], eb) * rng.standard_normal((frames, e["n"]))).astype(np.float32)
        fb, _, fi, ff = oracle.min_sum(H, y, "NMS", 0.8, 0.0, 50)
        qb, _, qi, qf = oracle.min_sum_fixed(H, y, "NMS_Q", 0.8, 0.0, 50, 0, 65536.0, 1 << 24, 1 << 24)
        both = (ff == 0) & (qf == 0)
        assert both.mean() > 0.75
        assert np.array_equal(fb[both], qb[both])  # under the reference stop rule both are the all-zero word
        wer_f, wer_q = (ff | fb.any(axis=1)).mean(), (qf | qb.any(axis=1)).mean()
        assert abs(wer_f - wer_q) < 3 * np.sqrt(wer_f * (1 - wer_f) / frames) + 0.01
        # first-iteration decisions (no accumulated rounding yet) agree on every frame
        f1 = oracle.min_sum(H, y, "NMS", 0.8, 0.0, 1, 2)
        q1 = oracle.min_sum_fixed(H, y, "NMS_Q", 0.8, 0.0, 1, 2, 65536.0, 1 << 24, 1 << 24)
        assert (f1[0] != q1[0]).mean() < 2e-4


@pytest.mark.parametrize("name,quant,floor_failed,floor_iter", [
    ("bch_15_7", (64.0, 255, 255), 0.99, 0.97), ("bch_31_16", (128.0, 1023, 255), 0.96, 0.88),
    ("bch_63_36", (64.0, 1023, 1023), 0.91, 0.79), ("bch_63_36", (8.0, 31, 31), 0.92, 0.67)])
def test_fixed_point_restatement_against_the_reference_on_golden_llrs(name, quant, floor_failed, floor_iter, catalogue,
                                                                      golden_codes):
    """SURVEY 7.2 step 8 cross-check.  The fixed-point min-sum (CCGPU_MS_Q / NMS_Q) has no reference implementation, so
    its restatement cannot be pinned bit for bit; what CAN be checked against the reference is that it is the SAME
    decoder up to quantisation: on the LLRs dumped from the reference (REF-FIXED outputs in tests/golden/) the integer
    decoder with a fine quantiser reproduces the float decoder's failure flag and iteration index on most frames.
    Measured disagreement (the golden frames sit at 0..4 dB, where 13..30 % of them fail and the iteration count of
    min-sum is chaotic): failure flag 0.2 % (15,7) / 3 % (31,16) / 6..8 % (63,36), iteration index among the frames both
    decode 1 % / 3..10 % / 19..31 %; the decided words of frames both decode are identical (the all-zero codeword).
    The floors below are those rates with a margin; the coarse default quantiser is included to show the trend."""
    g = load_golden("minsum_%s.npz" % name)
    H = golden_H(golden_codes, catalogue, name)
    n = catalogue[name]["n"]
    for v, (variant, alpha) in ((0, ("MS_Q", 1.0)), (1, ("NMS_Q", 0.8))):
        bits, _, it, failed = oracle.min_sum_fixed(H, g["y"], variant, alpha, 0.0, 50, 0, *quant)
        gf, gi = g["v%d.failed" % v], g["v%d.iter" % v]
        gb = np.unpackbits(g["v%d.bits" % v], axis=1)[:, :n]
        both = (gf == 0) & (failed == 0)
        assert (failed == gf).mean() >= floor_failed, (name, variant, (failed == gf).mean())
        assert (it[both] == gi[both]).mean() >= floor_iter, (name, variant, (it[both] == gi[both]).mean())
        assert np.array_equal(bits[both], gb[both])
        assert abs(int(failed.sum()) - int(gf.sum())) <= 0.2 * gf.sum() + 3  # same failure RATE within 20 %

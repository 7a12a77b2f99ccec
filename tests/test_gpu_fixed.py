"""GPU parity tests of the FIXED-POINT min-sum variants (CCGPU_MS_Q / NMS_Q / OMS_Q, include/ccgpu.h): the packed
two-frames-per-lane kernels (ms_cyclic_q.cuh, ms_cyclic_cta_q.cuh) against the integer restatement of the
reference's min_sum__ loop (oracle/ms_oracle.c, oracle_min_sum_fixed).  Parity is UNPINNED by the reference (it has no
integer decoder); the bar is bit-exact hard decisions, iteration indices, failure flags and integer totals."""
import zlib

import numpy as np
import pytest

import oracle

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    import channelcoding_b200 as cc
    c = cc.Context(0)
    yield c
    c.close()


def make_code(ctx, e):
    return ctx.bch(e["q"], **({"errors": e["cap_value"]} if e["cap_kind"] == 0 else {"dmin": e["cap_value"]}))


def assert_same_fixed(gpu, ref, what):
    gb, gL, gi, gf = gpu
    ob, oL, oi, of = ref
    assert np.array_equal(gf, of), what + ": failed flags"
    assert np.array_equal(gi.astype(np.uint32), oi.astype(np.uint32)), what + ": iteration index"
    assert np.array_equal(gb, ob), what + ": bits"
    if gL is not None:
        assert np.array_equal(gL, oL.astype(np.float32)), what + ": integer totals L"


# quantiser (scale, y_max, msg_max) per column weight so that w * fn_h(msg_max) + y_max <= 2048
QUANT = {"bch_15_7": [(8.0, 31, 31), (64.0, 255, 255), (3.0, 7, 5)], "bch_31_16": [(8.0, 31, 31), (32.0, 127, 127)],
         "bch_63_36": [(8.0, 31, 31), (16.0, 63, 63), (24.0, 100, 100)], "bch_127_64": [(8.0, 31, 31), (16.0, 63, 63)],
         "bch_255_131": [(8.0, 31, 29), (4.0, 15, 15)], "bch_63_45": [(8.0, 31, 31)], "bch_127_106": [(8.0, 31, 31)],
         "bch_31_26": [(8.0, 31, 31)], "bch_63_57": [(8.0, 31, 31)]}


@pytest.mark.parametrize("name,frames,ebnos", [("bch_15_7", 4001, (0.0, 2.0, 5.0)), ("bch_31_16", 2001, (1.0, 4.0)),
                                                ("bch_63_36", 1501, (1.0, 3.0, 5.0)), ("bch_63_45", 301, (3.0,)),
                                                ("bch_31_26", 1500, (4.0,)), ("bch_63_57", 400, (5.0,)),
                                                ("bch_127_64", 61, (3.5,)), ("bch_127_106", 60, (5.0,)),
                                                ("bch_255_131", 25, (5.5,))])
def test_fixed_vs_restatement(ctx, name, frames, ebnos, catalogue):
    """fresh seeded noise (odd frame counts: the last lane slot stays empty), three variants, all stop rules,
    several quantisers; zero / saturating / non-finite inputs included"""
    e = catalogue[name]
    code = make_code(ctx, e)
    assert code.kernel == 1
    H = code.H()
    rng = np.random.default_rng(zlib.crc32(name.encode()) + 1)
    for eb in ebnos:
        y = (1 + oracle.sigma(e["rate"], eb) * rng.standard_normal((frames, e["n"]))).astype(np.float32)
        y[0] = 0.0
        y[1] = -1.0
        y[2, ::3] = np.float32(np.nan)
        y[3, ::2] = np.float32(np.inf)
        y[3, 1::4] = np.float32(-np.inf)
        y[4] *= 100.0          # everything saturates
        y[5] = 0.0625          # exact rounding ties of the quantiser at scale 8
        y[6] = -0.1875
        for quant in QUANT[name]:
            for variant, alpha, beta, mi, stop in (("MS_Q", 1, 0, 50, 0), ("NMS_Q", 0.8, 0, 50, 0), ("OMS_Q", 1, 0.3, 20, 0),
                                                   ("NMS_Q", 0.8, 0, 50, 1), ("MS_Q", 1, 0, 4, 2), ("OMS_Q", 1, 0.15, 12, 1),
                                                   ("NMS_Q", 0.915, 0, 30, 0), ("NMS_Q", 0.5, 0, 10, 1)):
                sel = slice(0, frames if e["n"] <= 63 else max(9, frames // 4))
                gpu = code.decode(y[sel], variant, alpha, beta, mi, stop, quant=quant)
                ref = oracle.min_sum_fixed(H, y[sel], variant, alpha, beta, mi, stop, *quant)
                assert_same_fixed(gpu, ref, "%s %s stop=%d ebno=%g quant=%s" % (name, variant, stop, eb, quant))


@pytest.mark.parametrize("frames", [0, 1, 2, 3, 7, 8, 9, 257])
def test_fixed_ragged_batches(ctx, frames, catalogue):
    """every small batch size: both slots of a lane, one of them, none"""
    for name in ("bch_15_7", "bch_63_36", "bch_255_131"):
        e = catalogue[name]
        code = make_code(ctx, e)
        rng = np.random.default_rng(frames)
        y = (1 + 0.8 * rng.standard_normal((frames, e["n"]))).astype(np.float32)
        q = QUANT[name][0]
        gpu = code.decode(y, "NMS_Q", 0.8, 0.0, 10, 0, quant=q)
        if frames == 0:
            assert gpu[0].shape == (0, e["n"])
            continue
        assert_same_fixed(gpu, oracle.min_sum_fixed(code.H(), y, "NMS_Q", 0.8, 0.0, 10, 0, *q), "%s %d frames" % (name, frames))


def test_fixed_nonzero_codewords(ctx, catalogue):
    """random codewords over BPSK with the GF(2) stop rule: the decoder must not depend on the all-zero word"""
    for name in ("bch_31_16", "bch_63_36", "bch_127_64"):
        e = catalogue[name]
        code = make_code(ctx, e)
        rng = np.random.default_rng(11)
        frames = 600 if e["n"] <= 63 else 40
        msgs = rng.integers(0, 2, size=(frames, e["l"])).astype(np.uint8)
        words = code.encode(msgs)
        y = ((1.0 - 2.0 * words) + oracle.sigma(e["rate"], 4.5) * rng.standard_normal(words.shape)).astype(np.float32)
        gpu = code.decode(y, "NMS_Q", 0.8, 0.0, 30, 1)
        ref = oracle.min_sum_fixed(code.H(), y, "NMS_Q", 0.8, 0.0, 30, 1)
        assert_same_fixed(gpu, ref, name)
        ok = gpu[3] == 0
        assert ok.mean() > (0.5 if e["n"] <= 63 else 0.2) and (gpu[0][ok] == words[ok]).all(axis=1).mean() > 0.95


def test_fixed_redundant_rows(ctx, catalogue):
    """redundant wrap-around H (column weight grows to the row weight x rows / n): quantiser chosen inside the bound"""
    for name, rows, quant in (("bch_15_7", 15, (8.0, 31, 31)), ("bch_63_36", 63, (8.0, 31, 31)), ("bch_63_36", 40, (8.0, 31, 31)),
                              ("bch_127_64", 127, (8.0, 31, 31)), ("bch_255_131", 255, (4.0, 15, 14))):
        e = catalogue[name]
        code = make_code(ctx, e)
        code.set_rows(rows)
        H = oracle.Code(0, e["q"], e["t"]).H(rows)
        rng = np.random.default_rng(rows)
        frames = 301 if e["n"] <= 63 else (25 if e["n"] <= 127 else 7)
        y = (1 + oracle.sigma(e["rate"], 3.0) * rng.standard_normal((frames, e["n"]))).astype(np.float32)
        for variant, alpha, stop in (("NMS_Q", 0.8, 1), ("MS_Q", 1.0, 0)):
            assert_same_fixed(code.decode(y, variant, alpha, 0.0, 15, stop, quant=quant),
                              oracle.min_sum_fixed(H, y, variant, alpha, 0.0, 15, stop, *quant), "%s %d rows %s" % (name, rows, variant))


def test_fixed_bound_is_enforced(ctx, catalogue):
    """parameter sets whose column sums could leave the exactly representable range are rejected, not mis-decoded"""
    import channelcoding_b200 as cc
    code = make_code(ctx, catalogue["bch_63_36"])
    y = np.ones((4, 63), np.float32)
    wcol = int(code.H().sum(axis=0).max())  # 12 for the 27-row H(), 18 for the redundant 63-row matrix
    bad = (2048 - 31) // wcol + 1
    with pytest.raises(cc.CcgpuError) as ei:
        code.decode(y, "MS_Q", quant=(8.0, 31, bad))
    assert ei.value.code == cc._lib.ERR_UNSUPPORTED and "2048" in str(ei.value)
    code.decode(y, "MS_Q", quant=(8.0, 31, bad - 1))
    code.decode(y, "NMS_Q", 0.8, quant=(8.0, 31, bad))  # fn_h(bad) = rne(0.8 bad) fits
    code.set_rows(63)
    with pytest.raises(cc.CcgpuError):
        code.decode(y, "MS_Q", quant=(8.0, 31, bad - 1))  # column weight 18 now
    code.set_rows(27)
    with pytest.raises(cc.CcgpuError):
        code.decode(y, "NMS_Q", 1.5)
    g = ctx.from_dense(code.H()[np.random.default_rng(0).permutation(27)], 36 / 63)  # general H: CSR kernel, float only
    with pytest.raises(cc.CcgpuError):
        g.decode(y, "MS_Q")


def test_fixed_fused_point_equals_streaming(ctx, catalogue):
    """ccgpu_awgn_point with a fixed-point variant == channel kernel + ccgpu_decode_llr on the same frames"""
    for name, frames in (("bch_63_36", 50001), ("bch_15_7", 100003), ("bch_255_131", 1501)):
        e = catalogue[name]
        code = make_code(ctx, e)
        q = QUANT[name][0]
        for eb in (3.0, 6.0):
            c = code.awgn_point(eb, frames, "NMS_Q", 0.8, seed=5, point=3, frame0=1000, quant=q)
            y = ctx.awgn_llr(e["n"], np.float32(oracle.sigma(e["rate"], eb)), seed=5, point=3, frame0=1000, frames=frames)
            bits, _, it, failed = code.decode(y, "NMS_Q", 0.8, quant=q, want_L=False)
            assert c["frames"] == frames
            assert c["failures"] == int(failed.sum())
            assert c["frame_errors"] == int(((failed == 1) | bits.any(axis=1)).sum())
            assert c["bit_errors"] == int(bits.sum())
            assert c["iterations"] == int(np.where(failed == 1, 50, it.astype(np.int64) + 1).sum())


def test_fixed_bitflip_matches_restatement(ctx, catalogue):
    """bitflip_simulation inputs (x = +-1) through the fixed-point decoder: counts equal the restatement's"""
    import itertools
    e = catalogue["bch_31_16"]
    code = make_code(ctx, e)
    H = code.H()
    for w in (1, 2, 3):
        pats = np.ones((math_comb(31, w), 31), np.float32)
        for i, pos in enumerate(itertools.combinations(range(31), w)):
            pats[i, list(pos)] = -1.0
        for variant, alpha in (("MS_Q", 1.0), ("NMS_Q", 0.8)):
            c = code.bitflip_point(w, variant, alpha)
            b, _, _, f = oracle.min_sum_fixed(H, pats, variant, alpha)
            assert c["frames"] == len(pats)
            assert c["frame_errors"] == int(((f == 1) | b.any(axis=1)).sum()), (w, variant)


def math_comb(n, k):
    import math
    return math.comb(n, k)


def test_fixed_wer_close_to_float(ctx, catalogue):
    """statistical leg: on the SAME noise (same Philox stream) the fixed-point decoder with a fine quantiser has the
    float decoder's WER within its 95 % interval; the default 6-bit quantiser loses less than 20 % relative WER"""
    e = catalogue["bch_15_7"]
    code = make_code(ctx, e)
    frames = 2_000_000
    for eb in (2.0, 4.0):
        f = code.awgn_point(eb, frames, "MS", seed=1, point=7)
        fine = code.awgn_point(eb, frames, "MS_Q", seed=1, point=7, quant=(64.0, 255, 255))
        coarse = code.awgn_point(eb, frames, "MS_Q", seed=1, point=7)
        pf, pq, pc = (c["frame_errors"] / frames for c in (f, fine, coarse))
        half = 1.96 * np.sqrt(pf * (1 - pf) / frames) * np.sqrt(2.0)
        assert abs(pq - pf) < half + 0.01 * pf, (eb, pf, pq)
        assert pc < 1.2 * pf, (eb, pf, pc)
    e = catalogue["bch_63_36"]
    code = make_code(ctx, e)
    f = code.awgn_point(4.0, frames, "NMS", 0.8, seed=1, point=8)
    c = code.awgn_point(4.0, frames, "NMS_Q", 0.8, seed=1, point=8, quant=(16.0, 63, 63))
    assert c["frame_errors"] < 1.15 * f["frame_errors"], (f, c)


def test_fixed_mbbp(ctx, catalogue):
    """multiple-bases decoding accepts the fixed-point variants: one base with shift 0 equals the plain call"""
    e = catalogue["bch_63_36"]
    code = make_code(ctx, e)
    rng = np.random.default_rng(2)
    y = (1 + 0.7 * rng.standard_normal((500, 63))).astype(np.float32)
    plain = code.decode(y, "NMS_Q", 0.8)
    # decode_mbbp has no quant argument: the defaults are the plain call's defaults
    b, L, it, f, ch = code.decode_mbbp(y, [0], "NMS_Q", 0.8)
    assert np.array_equal(b, plain[0]) and np.array_equal(it, plain[2]) and np.array_equal(f, plain[3])


@pytest.mark.parametrize("name,ebno", [("bch_15_7", 6.0), ("bch_31_16", 6.5), ("bch_63_36", 7.0), ("bch_127_64", 8.0),
                                       ("bch_255_131", 9.0)])
def test_fixed_all_positive_shortcut_is_exact(ctx, name, ebno, catalogue):
    """fresh frames whose QUANTISED channel values are all positive are retired inside the refill loop (no iteration
    executed) when the totals are not requested: identical bits / iteration index / failure flag / counters to the full
    path (option quick = 0, and L requested) and to the restatement; zeros after quantisation, NaNs and ties never take
    the shortcut wrongly"""
    import os
    e = catalogue[name]
    code = make_code(ctx, e)
    n = e["n"]
    rng = np.random.default_rng(zlib.crc32(("quickq" + name).encode()))
    frames = 6001 if n < 255 else 1501
    y = (1 + oracle.sigma(e["rate"], ebno) * rng.standard_normal((frames, n))).astype(np.float32)
    y[0] = 1.0
    y[1, 3] = 0.05           # positive, but quantises to 0 at scale 8
    y[2] = np.inf
    y[3, 5] = np.nan         # quantises to 0
    y[4] = 0.0625            # tie 0.5 -> 0
    y[5, n - 1] = -0.0
    q = QUANT[name][0]
    yi = np.clip(np.rint(np.nan_to_num(y * np.float32(q[0]), nan=0.0)), -q[1], q[1])
    allpos = (yi > 0).all(axis=1)
    assert 0.15 < allpos.mean() < 0.98
    for variant, alpha, beta, mi, stop in (("MS_Q", 1, 0, 50, 0), ("NMS_Q", 0.8, 0, 50, 1), ("OMS_Q", 1, 0.3, 20, 0),
                                           ("NMS_Q", 0.8, 0, 1, 0), ("MS_Q", 1, 0, 5, 2)):
        full = code.decode(y, variant, alpha, beta, mi, stop, want_L=True, quant=q)
        try:
            ctx.set_option("quick", 1)
            fast = code.decode(y, variant, alpha, beta, mi, stop, want_L=False, quant=q)
            ctx.set_option("quick", 0)
            slow = code.decode(y, variant, alpha, beta, mi, stop, want_L=False, quant=q)
        finally:
            ctx.set_option("quick", -1)
        what = "%s %s stop=%d" % (name, variant, stop)
        for other in (fast, slow):
            assert np.array_equal(other[0], full[0]) and np.array_equal(other[2], full[2]) and np.array_equal(other[3], full[3]), what
        if stop != 2:
            assert not fast[0][allpos].any() and not fast[2][allpos].any() and not fast[3][allpos].any(), what
    hi = 1500 if n < 255 else 60
    ref = oracle.min_sum_fixed(code.H(), y[:hi], "NMS_Q", 0.8, 0.0, 50, 0, *q)
    ctx.set_option("quick", 1)
    try:
        fast = code.decode(y[:hi], "NMS_Q", 0.8, 0.0, 50, 0, want_L=False, quant=q)
        c1 = code.awgn_point(ebno, 300001 if n < 255 else 60001, "NMS_Q", 0.8, seed=3, point=1, quant=q)
        ctx.set_option("quick", 0)
        c0 = code.awgn_point(ebno, 300001 if n < 255 else 60001, "NMS_Q", 0.8, seed=3, point=1, quant=q)
    finally:
        ctx.set_option("quick", -1)
    assert np.array_equal(fast[0], ref[0]) and np.array_equal(fast[2].astype(np.uint32), ref[2]) and np.array_equal(fast[3], ref[3])
    assert c0 == c1

"""GPU parity tests of the lane-per-frame kernel (csrc/ms_cyclic_lane.cuh, the n = 15 and n = 31 codes): bit-exact against the
oracle (pinned to the reference) and against the warp kernel it replaces (option "lane" = 0), through the C ABI."""
import zlib

import numpy as np
import pytest

import oracle
from conftest import golden_H, load_golden, VARIANT_PARAMS

pytestmark = pytest.mark.gpu

NAMES = ["bch_15_7", "bch_15_5", "bch_15_7_dmin5", "bch_15_7_dmin6", "bch_31_16", "bch_31_26", "bch_31_21", "bch_31_11"]


@pytest.fixture(scope="module")
def ctx():
    import channelcoding_b200 as cc
    c = cc.Context(0)
    yield c
    c.close()


def make_code(ctx, e):
    return ctx.bch(e["q"], **({"errors": e["cap_value"]} if e["cap_kind"] == 0 else {"dmin": e["cap_value"]}))


def same(a, b, what):
    assert np.array_equal(a[3], b[3]), what + ": failed flags"
    assert np.array_equal(np.asarray(a[2]).astype(np.uint32), np.asarray(b[2]).astype(np.uint32)), what + ": iteration index"
    assert np.array_equal(a[0], b[0]), what + ": bits"
    if a[1] is not None and b[1] is not None:
        assert np.array_equal(a[1].view(np.uint32), b[1].view(np.uint32)), what + ": totals L (bit pattern)"


@pytest.mark.parametrize("name", NAMES)
def test_lane_kernel_vs_oracle_and_warp_kernel(ctx, name, catalogue):
    """every min-sum variant x stop rule on seeded noise at 0 .. 7 dB, ragged frame counts (1, 31, 33, 4099 -- not a
    multiple of the warp, the FIFO pass or the CTA), zeros / ties / infinities / NaNs in the input: bits, totals L (bit
    pattern), iteration index and failure flag equal the oracle's and the warp kernel's"""
    e = catalogue[name]
    code = make_code(ctx, e)
    n = e["n"]
    H = code.H()
    rng = np.random.default_rng(zlib.crc32(("lane" + name).encode()))
    for frames, eb in ((1, 3.0), (31, 1.0), (33, 5.0), (4099, 0.0), (4099, 3.0), (20011, 6.0)):
        y = (1 + oracle.sigma(e["rate"], eb) * rng.standard_normal((frames, n))).astype(np.float32)
        if frames > 100:
            y[3, 2] = 0.0
            y[4] = 1.0                      # every magnitude ties
            y[5, :] = -1.0
            y[6, 1] = np.inf
            y[7, 0] = -np.inf
            y[8, 4] = np.nan
            y[9] = np.float32(1e-45)
            y[10, n - 1] = -0.0
        finite = np.isfinite(y).all(axis=1)
        for variant, alpha, beta, mi, stop in (("MS", 1, 0, 50, 0), ("NMS", 0.8, 0, 50, 0), ("OMS", 1, 0.3, 20, 0), ("NMS", 0.8, 0, 50, 1),
                                               ("2DNMS", 0.9, 0.8, 25, 0), ("2DNMS", 0.9, 0.8, 25, 1), ("MS", 1, 0, 7, 2), ("NMS", 0.8, 0, 1, 0)):
            what = "%s %s stop=%d %d frames %.1f dB" % (name, variant, stop, frames, eb)
            ctx.set_option("lane", 1)
            try:
                lane = code.decode(y, variant, alpha, beta, mi, stop, want_L=True)
                ctx.set_option("lane", 0)
                warp = code.decode(y, variant, alpha, beta, mi, stop, want_L=True)
            finally:
                ctx.set_option("lane", -1)
            same(lane, warp, what + " (lane vs warp kernel)")
            ref = oracle.min_sum(H, y[finite], variant, alpha, beta, mi, stop)
            same(tuple(None if a is None else a[finite] for a in lane), ref, what + " (lane kernel vs oracle)")


def test_lane_kernel_golden(ctx, catalogue, golden_codes):
    """the reference's own dumped LLRs of BCH(15,7): all variants bit-exact (this is test_decode_golden, asserted here to
    run on the lane kernel)"""
    name = "bch_15_7"
    g = load_golden("minsum_%s.npz" % name)
    e = catalogue[name]
    code = make_code(ctx, e)
    n = e["n"]
    before = ctx.kernel_launches
    for v, (variant, alpha, beta, max_iter) in VARIANT_PARAMS.items():
        if "v%d.iter" % v not in g or variant.startswith("SCMS"):
            continue
        bits, L, it, failed = code.decode(g["y"], variant, alpha, beta, max_iter)
        assert np.array_equal(failed, g["v%d.failed" % v]) and np.array_equal(it, g["v%d.iter" % v]), (name, v)
        ok = failed == 0
        assert np.array_equal(bits[ok], np.unpackbits(g["v%d.bits" % v], axis=1)[:, :n][ok]), (name, v)
        if "v%d.L" % v in g:
            assert np.array_equal(L[ok].view(np.uint32), g["v%d.L" % v][ok].view(np.uint32)), (name, v)
    assert ctx.kernel_launches > before


@pytest.mark.parametrize("name", ["bch_15_7", "bch_15_5", "bch_31_16", "bch_31_26"])
def test_lane_kernel_compact_outputs_and_counters(ctx, name, catalogue):
    """compact output layout, and the fused Monte-Carlo point: identical counters on the lane kernel and the warp kernel
    (the noise is keyed by the frame index), for ragged frame counts and frame offsets; sum-product included"""
    e = catalogue[name]
    code = make_code(ctx, e)
    n = e["n"]
    rng = np.random.default_rng(5)
    y = (1 + oracle.sigma(e["rate"], 3.0) * rng.standard_normal((5003, n))).astype(np.float32)
    bits, _, it, failed = code.decode(y, "NMS", 0.8, 0.0, 50, 0, want_L=False)
    packed, status = code.decode_packed(y, "NMS", 0.8, 0.0, 50, 0)
    assert np.array_equal(status == 255, failed == 1) and np.array_equal(status[failed == 0], it[failed == 0])
    assert np.array_equal(((packed[:, 0][:, None] >> np.arange(n)) & 1).astype(np.uint8), bits)
    for eb, frames, variant, alpha, beta, mi, stop, f0 in ((1.0, 100003, "MS", 1.0, 0.0, 50, 0, 0), (3.0, 400001, "NMS", 0.8, 0.0, 50, 0, 17),
                                                             (6.0, 1000003, "NMS", 0.8, 0.0, 50, 1, 1 << 33), (9.0, 2000000, "OMS", 1.0, 0.2, 20, 0, 0),
                                                             (3.0, 7, "MS", 1.0, 0.0, 50, 0, 0), (3.0, 300001, "2DNMS", 0.9, 0.8, 25, 0, 3),
                                                             (3.0, 300001, "SPA", 1.0, 0.0, 50, 1, 0), (6.0, 300001, "SPA", 1.0, 0.0, 50, 1, 5)):
        res = []
        for lane in (1, 0):
            ctx.set_option("lane", lane)
            try:
                res.append(code.awgn_point(eb, frames, variant, alpha, beta, mi, stop, seed=4, point=2, frame0=f0))
            finally:
                ctx.set_option("lane", -1)
        assert res[0] == res[1] and res[0]["frames"] == frames, (name, variant, eb, res)


def test_lane_kernel_sum_product_matches_the_warp_kernel(ctx, catalogue):
    """sum-product (extension): the lane kernel evaluates the same expressions in the same order as the warp kernel
    (prefix x suffix products, tanhf / atanhf, clamp), so bits, iteration index, failure flag AND the totals' bit patterns
    agree; the tolerance against the float64 yardstick is tests/test_gpu_parity.py's business"""
    e = catalogue["bch_15_7"]
    code = make_code(ctx, e)
    rng = np.random.default_rng(77)
    for eb in (1.0, 3.0, 6.0):
        sg = oracle.sigma(e["rate"], eb)
        y = ((1 + sg * rng.standard_normal((30011, e["n"]))) * (2.0 / sg ** 2)).astype(np.float32)
        out = []
        for lane in (1, 0):
            ctx.set_option("lane", lane)
            try:
                out.append(code.decode(y, "SPA", 1.0, 0.0, 50, 1, want_L=True))
            finally:
                ctx.set_option("lane", -1)
        same(out[0], out[1], "SPA %.1f dB" % eb)

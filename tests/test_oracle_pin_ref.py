"""The C restatement (oracle/) against the reference itself (oracle/_ref/libccref.so) on fresh
random inputs.  Skipped where oracle/_ref was not built (needs /root/reference at build time)."""
import numpy as np
import pytest

import ccref
import oracle
from ccref import ALG_EUKLID, VARIANT_PARAMS

pytestmark = pytest.mark.skipif(not ccref.available(), reason="oracle/_ref/libccref.so not built")


@pytest.fixture(scope="module")
def ref():
    return ccref.Ref()


@pytest.mark.parametrize("q,kind,val,t,ebnos,frames", [(4, 0, 2, 2, (1, 4), 150), (5, 1, 7, 3, (3,), 100),
                                                       (6, 0, 5, 5, (2, 5), 40), (7, 0, 10, 10, (4,), 6)])
def test_min_sum_pin(ref, q, kind, val, t, ebnos, frames):
    rng = np.random.default_rng(q * 100 + t)
    c = oracle.Code(0, q, t)
    H = c.H()
    assert np.array_equal(H, ref.H(0, q, kind, val))
    for eb in ebnos:
        y = (1 + oracle.sigma(c.rate, eb) * rng.standard_normal((frames, c.n))).astype(np.float32)
        y[0, :] = 1
        y[1, :] = 0
        y[2, :] = 1; y[2, 3] = 0; y[2, 7] = -1
        for v, (name, a, b, mi) in VARIANT_PARAMS.items():
            rb, rL, rit, rf = ref.min_sum(v, H, y)
            ob, oL, oit, of = oracle.min_sum(H, y, name, a, b, mi)
            assert np.array_equal(rf, of) and np.array_equal(rit, oit), (q, eb, v)
            ok = rf == 0
            assert np.array_equal(rb[ok], ob[ok]), (q, eb, v)
            assert np.array_equal(rL[ok].view(np.uint32), oL[ok].view(np.uint32)), (q, eb, v)


def test_ref_head_is_one_iteration(ref):
    """HEAD (matrix.h:50 bug) == REF-FIXED truncated to one iteration, never failing."""
    if not ccref.available(head=True):
        pytest.skip("libccref_head.so not built")
    head = ccref.Ref(head=True)
    rng = np.random.default_rng(5)
    c = oracle.Code(0, 5, 3)
    y = (1 + 0.8 * rng.standard_normal((200, c.n))).astype(np.float32)
    hb, hL, hit, hf = head.min_sum(1, c.H(), y)
    ob, oL, oit, of = oracle.min_sum(c.H(), y, "NMS", 0.8, 0.0, 1, oracle.STOP_NONE)
    assert not hf.any() and not hit.any() and not of.any()
    assert np.array_equal(hb, ob) and np.array_equal(hL.view(np.uint32), oL.view(np.uint32))


@pytest.mark.parametrize("fam,q,kind,val,t,count", [(1, 8, 0, 16, 16, 120), (1, 4, 0, 3, 3, 500), (1, 3, 0, 2, 2, 500),
                                                    (0, 4, 0, 2, 2, 500), (0, 6, 0, 5, 5, 300), (0, 8, 0, 18, 18, 40)])
def test_hard_pin(ref, fam, q, kind, val, t, count):
    rng = np.random.default_rng(fam * 1000 + q * 10 + t)
    c = oracle.Code(fam, q, t)
    msgs = rng.integers(0, (1 << q) if fam == 1 else 2, size=(count, c.l)).astype(np.uint8)
    words = ref.encode(fam, q, kind, val, msgs)
    assert np.array_equal(words, c.encode(msgs))
    bad = words.copy()
    for i in range(count):
        ne = rng.integers(0, t + 4)
        pos = rng.choice(c.n, ne, replace=False)
        bad[i, pos] ^= (rng.integers(1, 1 << q, size=ne).astype(np.uint8) if fam == 1 else 1)
    ro, rs = ref.hard_correct(fam, q, kind, val, ALG_EUKLID, bad)
    oo, on, os_ = c.hard_correct(bad)
    assert np.array_equal(rs, os_)
    assert np.array_equal(ro[rs == 0], oo[rs == 0])


def test_erasures_pin(ref):
    """Euklid with erasures (hard_decision.h:171-172) on RS(15,9): e errors + f erasures, 2e+f <= 2t."""
    rng = np.random.default_rng(11)
    fam, q, kind, val, t = 1, 4, 0, 3, 3
    c = oracle.Code(fam, q, t)
    for trial in range(60):
        msg = rng.integers(0, 16, size=(1, c.l)).astype(np.uint8)
        word = c.encode(msg)
        f = int(rng.integers(1, 5))
        e = int(rng.integers(0, (2 * t - f) // 2 + 2))
        pos = rng.choice(c.n, f + e, replace=False)
        bad = word.copy()
        bad[0, pos[:f]] = 0
        bad[0, pos[f:]] ^= rng.integers(1, 16, size=e).astype(np.uint8)
        er = sorted(int(p) for p in pos[:f])
        ro, rs = ref.hard_correct(fam, q, kind, val, ALG_EUKLID, bad, er)
        oo, on, os_ = c.hard_correct(bad, er)
        assert rs[0] == os_[0], (trial, er, e)
        if rs[0] == 0:
            assert np.array_equal(ro, oo)

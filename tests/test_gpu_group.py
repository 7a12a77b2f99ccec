"""ccgpu_group through the Python binding: sharded Monte-Carlo points and batched decodes equal the one-device results.
The members may all sit on device 0 (the logic is the same); with several GPUs the real devices are used as well."""
import numpy as np
import pytest

import oracle

pytestmark = pytest.mark.gpu


def device_sets():
    import torch
    sets = [[0, 0, 0]]
    if torch.cuda.device_count() >= 2:
        sets.append(list(range(torch.cuda.device_count())))
    return sets


@pytest.mark.parametrize("which", [0, 1])
def test_group_points_equal_single_device(which):
    import channelcoding_b200 as cc
    sets = device_sets()
    if which >= len(sets):
        pytest.skip("one GPU only")
    ctx = cc.Context(0)
    g = cc.Group(sets[which])
    try:
        one, many = ctx.bch(6, errors=5), g.bch(6, errors=5)
        for frames in (1, 5, 16384, 100003, 700001):
            for variant, alpha, quant in (("NMS", 0.8, None), ("NMS_Q", 0.8, (8.0, 31, 31)), ("SCMS2", 1.0, None)):
                a = one.awgn_point(4.0, frames, variant, alpha, seed=9, point=2, frame0=123, quant=quant)
                b = many.awgn_point(4.0, frames, variant, alpha, seed=9, point=2, frame0=123, quant=quant)
                assert a == b, (frames, variant)
            assert one.awgn_point_hard(5.0, frames, seed=9, point=2) == many.awgn_point_hard(5.0, frames, seed=9, point=2)
        g.set_min_frames(1)  # force every member to take part even in tiny points
        assert one.awgn_point(3.0, 7, "MS", seed=1) == many.awgn_point(3.0, 7, "MS", seed=1)
        assert one.awgn_point(3.0, 2, "MS", seed=1) == many.awgn_point(3.0, 2, "MS", seed=1)
        g.set_min_frames(16384)
        sh = [0, 9, 20]
        assert one.awgn_point_mbbp(5.0, 60001, sh, "NMS", 0.8, seed=4) == many.awgn_point_mbbp(5.0, 60001, sh, "NMS", 0.8, seed=4)
        one31, many31 = ctx.bch(5, dmin=7), g.bch(5, dmin=7)
        for w in range(5):
            assert one31.bitflip_point(w, "MS") == many31.bitflip_point(w, "MS")
        assert many31.bitflip_point(3, "MS", first=1000, count=2000) == one31.bitflip_point(3, "MS", first=1000, count=2000)
        # batched decode, host buffers
        rng = np.random.default_rng(1)
        y = (1 + 0.7 * rng.standard_normal((50001, 63))).astype(np.float32)
        a, b = one.decode(y, "NMS", 0.8), many.decode(y, "NMS", 0.8)
        for x, z in zip(a, b):
            assert np.array_equal(x.view(np.uint32) if x.dtype == np.float32 else x, z.view(np.uint32) if z.dtype == np.float32 else z)
        rs1, rsn = ctx.rs(8, 16), g.rs(8, 16)
        words = rs1.encode(rng.integers(0, 256, size=(20001, rs1.l)).astype(np.uint8))
        words[::3, 5] ^= 77
        for x, z in zip(rs1.gf_decode(words), rsn.gf_decode(words)):
            assert np.array_equal(x, z)
        # errors surface with the member's text
        with pytest.raises(cc.CcgpuError):
            many.awgn_point(4.0, 100000, "MS_Q", quant=(8.0, 31, 1000))
    finally:
        g.close()
        ctx.close()

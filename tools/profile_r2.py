#!/usr/bin/env python
"""tools/profile_r2.py -- ONE launch of every shipped kernel flavour, in a fixed order, for Nsight Compute:

    ncu --set full --clock-control none --import-source on -o gpurun_out/r2_kernels python tools/profile_r2.py
    python tools/ncu_summary.py gpurun_out/r2_kernels.ncu-rep profiles/r2_kernels --manifest gpurun_out/r2_manifest.json

The manifest (written next to the report) names every launch: label, units (frames / words) and the average number of
decoder iterations, so that the summary can state per-unit figures.  Run without ncu it prints CUDA-event timings.
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import numpy as np
    import torch
    import channelcoding_b200 as cc
    ctx = cc.Context(0)
    ctx.use_torch_stream()
    manifest = []

    def ev(fn):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1)

    def decode_target(label, match, code, eb, frames, variant, alpha=0.8, quant=None, quick=-1, packed=False):
        y = torch.empty((frames, code.n), dtype=torch.float32, device="cuda")
        ctx.awgn_llr(code.n, np.float32(cc.sigma(code.rate, eb)), 0, 1, 0, frames, out=y)
        manifest.append({"label": "K1 awgn_llr n=%d (input of %s)" % (code.n, label), "match": "awgn_llr_kernel", "units": frames,
                         "unit": "frame", "bytes_per_unit": 4 * code.n})
        out = (torch.empty((frames, code.n), dtype=torch.uint8, device="cuda"), None,
               torch.empty(frames, dtype=torch.uint8, device="cuda"), torch.empty(frames, dtype=torch.uint8, device="cuda"))
        ctx.set_option("quick", quick)
        if packed:  # compact output layout: the 264 algorithmic bytes per frame of SURVEY 8(d)
            pk = (torch.empty((frames, (code.n + 31) // 32), dtype=torch.int32, device="cuda"), torch.empty(frames, dtype=torch.uint8, device="cuda"))
            ms = ev(lambda: code.decode_packed(y, variant, alpha, 0.0, 50, out=pk, quant=quant))
            it = torch.where(pk[1] == 255, torch.full_like(pk[1], 50).int(), pk[1].int() + 1).double().mean().item()
        else:
            ms = ev(lambda: code.decode(y, variant, alpha, 0.0, 50, out=out, want_L=False, quant=quant))
            it = torch.where(out[3] == 1, torch.full_like(out[2], 50).int(), out[2].int() + 1).double().mean().item()
        ctx.set_option("quick", -1)
        manifest.append({"label": label, "match": match, "units": frames, "unit": "frame", "avg_iterations": it,
                         "edges": code.edges, "bytes_per_unit": 4 * code.n + 4 * ((code.n + 31) // 32) + 4, "ms_no_ncu": ms})
        print("%-46s %8.3f ms  %.3e frames/s  %.2f it" % (label, ms, frames / ms * 1e3, it))

    def point_target(label, match, code, eb, frames, variant, alpha=0.8, quant=None, quick=-1, stop=0):
        cnt = torch.zeros(8, dtype=torch.int64, device="cuda")
        ctx.set_option("quick", quick)
        ms = ev(lambda: code.awgn_point(eb, frames, variant, alpha, 0.0, 50, stop, seed=1, point=2, out=cnt, quant=quant))
        ctx.set_option("quick", -1)
        c = cnt.cpu().numpy()
        manifest.append({"label": label, "match": match, "units": frames, "unit": "frame", "avg_iterations": float(c[3]) / frames,
                         "edges": code.edges, "bytes_per_unit": 0, "ms_no_ncu": ms})
        print("%-46s %8.3f ms  %.3e frames/s  %.2f it" % (label, ms, frames / ms * 1e3, float(c[3]) / frames))

    c63 = ctx.bch(6, errors=5)
    M = 1 << 20
    decode_target("K2 ms_cyclic BCH(63,36) NMS 4 dB resident", "ms_cyclic_kernel", c63, 4.0, M, "NMS")
    decode_target("K2 ms_cyclic BCH(63,36) NMS 4 dB resident, compact outputs", "ms_cyclic_kernel", c63, 4.0, M, "NMS", packed=True)
    decode_target("K2q ms_cyclic_q BCH(63,36) NMS_Q 4 dB resident", "ms_cyclic_q_kernel", c63, 4.0, M, "NMS_Q", quant=(8.0, 31, 31))
    point_target("K2 ms_cyclic BCH(63,36) NMS 4 dB fused", "ms_cyclic_kernel", c63, 4.0, M, "NMS")
    point_target("K2q ms_cyclic_q BCH(63,36) NMS_Q 4 dB fused", "ms_cyclic_q_kernel", c63, 4.0, M, "NMS_Q", quant=(8.0, 31, 31))
    point_target("K2 QUICK BCH(63,36) NMS 7 dB fused", "ms_cyclic_kernel", c63, 7.0, 4 * M, "NMS", quick=1)
    point_target("K2 QUICK screening BCH(63,36) NMS 10 dB fused", "ms_cyclic_kernel", c63, 10.0, 32 * M, "NMS", quick=1)
    point_target("K2q skip BCH(63,36) NMS_Q 8 dB fused", "ms_cyclic_q_kernel", c63, 8.0, 4 * M, "NMS_Q", quant=(8.0, 31, 31), quick=1)
    point_target("K2 SC BCH(63,36) SCMS2 4 dB fused", "ms_cyclic_kernel", c63, 4.0, M, "SCMS2", 1.0)
    c15 = ctx.bch(4, errors=2)
    point_target("K2s ms_cyclic_lane BCH(15,7) sum-product 3 dB fused", "ms_cyclic_lane_kernel", c15, 3.0, 4 * M, "SPA", 1.0, stop=1)
    point_target("K2s ms_cyclic_lane BCH(15,7) MS 3 dB fused", "ms_cyclic_lane_kernel", c15, 3.0, 16 * M, "MS", 1.0)
    ctx.set_option("lane", 0)
    point_target("K2 ms_cyclic BCH(15,7) MS 3 dB fused (warp kernel, option lane = 0)", "ms_cyclic_kernel", c15, 3.0, 8 * M, "MS", 1.0)
    ctx.set_option("lane", -1)
    c31 = ctx.bch(5, errors=3)
    point_target("K2s ms_cyclic_lane BCH(31,16) NMS 4 dB fused", "ms_cyclic_lane_kernel", c31, 4.0, 4 * M, "NMS")
    ctx.set_option("lane", 0)
    point_target("K2 ms_cyclic BCH(31,16) NMS 4 dB fused (warp kernel, option lane = 0)", "ms_cyclic_kernel", c31, 4.0, 4 * M, "NMS")
    ctx.set_option("lane", -1)
    c127 = ctx.bch(7, errors=10)
    point_target("K2 ms_cyclic BCH(127,64) NMS 5 dB fused", "ms_cyclic_kernel", c127, 5.0, M // 4, "NMS")
    point_target("K2q ms_cyclic_q BCH(127,64) NMS_Q 5 dB fused", "ms_cyclic_q_kernel", c127, 5.0, M // 4, "NMS_Q", quant=(8.0, 31, 31))
    c127.set_rows(127)
    point_target("K2c ms_cyclic_cta BCH(127,64) 127-row H NMS 5 dB", "ms_cyclic_cta_kernel", c127, 5.0, M // 8, "NMS")
    point_target("K2cq ms_cyclic_cta_q BCH(127,64) 127-row H NMS_Q 5 dB", "ms_cyclic_cta_q_kernel", c127, 5.0, M // 8, "NMS_Q",
                 quant=(8.0, 31, 31))
    c255 = ctx.bch(8, errors=18)
    point_target("K2c ms_cyclic_cta BCH(255,131) NMS 6 dB fused", "ms_cyclic_cta_kernel", c255, 6.0, M // 8, "NMS")
    point_target("K2cq ms_cyclic_cta_q BCH(255,131) NMS_Q 6 dB fused", "ms_cyclic_cta_q_kernel", c255, 6.0, M // 8, "NMS_Q",
                 quant=(8.0, 31, 29))
    point_target("K2c grouped BCH(255,131) NMS 11.5 dB fused", "ms_cyclic_cta_kernel", c255, 11.5, 16 * M, "NMS")
    point_target("K2cq grouped BCH(255,131) NMS_Q 11.5 dB fused", "ms_cyclic_cta_q_kernel", c255, 11.5, 16 * M, "NMS_Q", quant=(8.0, 31, 29))
    # K2g: the (63,45) checks in a non-cyclic row order -> CSR kernel
    c45 = ctx.bch(6, dmin=7)
    g = ctx.from_dense(c45.H()[np.random.default_rng(1).permutation(c45.h_rows)], c45.rate)
    point_target("K2g ms_csr (63,45) permuted H NMS 4 dB fused", "ms_csr_kernel", g, 4.0, M // 8, "NMS")
    # K4
    rs = ctx.rs(8, 16)
    rng = np.random.default_rng(5)
    words = rs.encode(rng.integers(0, 256, size=(4096, rs.l)).astype(np.uint8))
    bad = words.copy()
    for i in range(4096):
        ne = i % 18
        pos = rng.choice(255, ne, replace=False)
        bad[i, pos] ^= rng.integers(1, 256, size=ne).astype(np.uint8)
    count = M
    for label, src in (("K4 gf_decode RS(255,223) 0..17 errors", bad), ("K4 gf_decode RS(255,223) error free", words)):
        d = torch.from_numpy(src).cuda().repeat(count // 4096, 1).contiguous()
        out = (torch.empty_like(d), torch.empty(count, dtype=torch.uint8, device="cuda"), torch.empty(count, dtype=torch.uint8, device="cuda"))
        ms = ev(lambda: rs.gf_decode(d, out=out))
        manifest.append({"label": label, "match": "gf_decode_kernel", "units": count, "unit": "word", "bytes_per_unit": 511, "ms_no_ncu": ms})
        print("%-46s %8.3f ms  %.3e words/s  %.1f GB/s algorithmic" % (label, ms, count / ms * 1e3, count * 511 / ms / 1e6))
    os.makedirs(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out"), exist_ok=True)
    with open(os.environ.get("CCGPU_MANIFEST", "gpurun_out/r2_manifest.json"), "w") as f:
        json.dump(manifest, f, indent=1)
    ctx.close()


if __name__ == "__main__":
    main()

"""Host logic of the sharded sweep (channelcoding_b200/simulation.py), CPU only: the frame-range
partition, the sweep schedule / log format of simulation.c++:95-150, and the N > 1 path with two gloo
processes (the per-point all-reduce of the counters) using a deterministic stand-in for the kernel."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from channelcoding_b200 import simulation


def fake_point(ebno, point, frame0, frames):
    """deterministic per-frame outcome keyed by the GLOBAL frame index, like the Philox-keyed kernel"""
    idx = torch.arange(frame0, frame0 + frames, dtype=torch.int64)
    h = (idx * 2654435761 + point * 40503) % 1000
    err = h < int(1000 * 0.5 * 10 ** (-ebno / 4.0))
    its = 1 + (h % 7)
    c = torch.zeros(8, dtype=torch.int64)
    c[0] = frames
    c[1] = int(err.sum())
    c[2] = int((err * 3).sum())
    c[3] = int(its.sum())
    c[4] = int(err.sum())
    return c


def test_shard_range_partitions():
    for total in (0, 1, 7, 10000, 12345677):
        for world in (1, 2, 3, 8):
            parts = [simulation.shard_range(total, world, r) for r in range(world)]
            assert parts[0][0] == 0 and sum(c for _, c in parts) == total
            for (f0, c0), (f1, _) in zip(parts, parts[1:]):
                assert f0 + c0 == f1
            assert max(c for _, c in parts) - min(c for _, c in parts) <= 1


def test_schedule_and_log_format(tmp_path):
    pts = simulation.sweep_points(36 / 63)
    assert pts[0] == 1.5 and pts[-1] == 8.0 and len(pts) == 14          # (63,36) starts at 1.5 dB
    assert simulation.sweep_points(7 / 15)[0] == 1.0
    assert simulation.LOG_HEADER == "   ebno                   wer"
    assert simulation.format_log_line(1.5, 0.25) == "    1.5 2.500000000000000e-01"
    assert simulation.format_log_line(8.0, 0.0) == "      8 0.000000000000000e+00"
    res = simulation.awgn_sweep(fake_point, "(63, 36, 11)-NMS", 36 / 63, log_dir=str(tmp_path))
    lines = open(tmp_path / "(63, 36, 11)-NMS.log").read().splitlines()
    assert lines[0] == simulation.LOG_HEADER and len(lines) == 15
    assert res[0]["frames"] == 10000                                     # 5e3 / 0.5 (simulation.c++:91-93,109)
    for prev, cur in zip(res, res[1:]):
        expect = 1000000 if prev["wer"] == 0 else min(1000000, int(5e3 / prev["wer"]))
        assert cur["frames"] == expect
    with pytest.raises(RuntimeError):                                    # never overwrite a log (simulation.c++:72-81)
        simulation.awgn_sweep(fake_point, "(63, 36, 11)-NMS", 36 / 63, log_dir=str(tmp_path))


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    res = simulation.awgn_sweep(fake_point, "x", 36 / 63, dist=dist)
    q.put((rank, res))
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_equal_one_rank():
    """world_size 2 over gloo: the merged counters of every point equal the single-process run"""
    single = simulation.awgn_sweep(fake_point, "x", 36 / 63)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert got[0] == single and got[1] == single

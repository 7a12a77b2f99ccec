#!/usr/bin/env python
"""Summarise an Nsight Compute report (.ncu-rep) into a small text + json file for profiles/.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r1_ms_cyclic [--frames-per-launch N]
    python tools/ncu_summary.py gpurun_out/r2_kernels.ncu-rep profiles/r2_kernels --manifest gpurun_out/r2_manifest.json

With a manifest (tools/profile_r2.py) every launch gets its label, per-unit figures (warp instructions, shared-memory
wavefronts and DRAM bytes per frame / word) and the pipe fractions; every record is stamped with the hash of
channelcoding_b200/csrc/ so that a reader (bench.py) can tell whether the capture describes the library it runs.
"""
import csv
import hashlib
import io
import json
import os
import subprocess
import sys


def csrc_hash():
    """sha256 over the kernel / host sources of the library (names and contents, sorted)"""
    root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "channelcoding_b200", "csrc")
    h = hashlib.sha256()
    for name in sorted(os.listdir(root)):
        if name.endswith((".cu", ".cuh", ".h", ".hpp", ".cc")):
            h.update(name.encode())
            with open(os.path.join(root, name), "rb") as f:
                h.update(f.read())
    return h.hexdigest()[:16]

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__waves_per_multiprocessor",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__warps_active.avg.per_cycle_active",
    "smsp__warps_eligible.avg.per_cycle_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__cycles_active.avg", "sm__cycles_elapsed.max",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_active", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
]
UNIT_SCALE = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0, "Tbyte": 1e12}


def main():
    rep, out = sys.argv[1], sys.argv[2]
    fpl = int(sys.argv[sys.argv.index("--frames-per-launch") + 1]) if "--frames-per-launch" in sys.argv else None
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    res = []
    manifest = None
    if "--manifest" in sys.argv:
        with open(sys.argv[sys.argv.index("--manifest") + 1]) as f:
            manifest = json.load(f)
    stamp = csrc_hash()
    mi = 0
    for r in rows[2:]:
        d = {"kernel": r[hdr.index("Kernel Name")], "csrc_sha": stamp}
        if manifest is not None:
            # launches appear in the manifest's order; helper kernels (memsets, counters) in between are skipped
            if mi >= len(manifest) or manifest[mi]["match"] not in d["kernel"]:
                continue
            d.update({k: v for k, v in manifest[mi].items() if k != "match"})
            mi += 1
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                try:
                    v = float(r[i].replace(",", ""))
                except ValueError:
                    continue
                if units[i] in UNIT_SCALE:
                    v *= UNIT_SCALE[units[i]]
                    d[k + " [byte]"] = v
                else:
                    d[k + (" [%s]" % units[i] if units[i] else "")] = v
        if manifest is not None and d.get("units"):
            u = float(d["units"])
            ins = [v for k, v in d.items() if k.startswith("smsp__inst_executed.sum")]
            wf = [v for k, v in d.items() if k.startswith("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum")]
            rd, wr = d.get("dram__bytes_read.sum [byte]"), d.get("dram__bytes_write.sum [byte]")
            if ins:
                d["warp_instructions_per_unit"] = ins[0] / u
            if wf:
                d["smem_wavefronts_per_unit"] = wf[0] / u
            if rd is not None and wr is not None:
                d["dram_bytes_per_unit"] = (rd + wr) / u
            if d.get("avg_iterations") and ins:
                d["warp_instructions_per_frame_iteration"] = ins[0] / u / d["avg_iterations"]
                if wf:
                    d["smem_wavefronts_per_frame_iteration"] = wf[0] / u / d["avg_iterations"]
        if fpl:
            rd, wr = d.get("dram__bytes_read.sum [byte]"), d.get("dram__bytes_write.sum [byte]")
            if rd is not None and wr is not None:
                d["dram_bytes_per_frame"] = (rd + wr) / fpl
                d["frames_per_launch"] = fpl
        res.append(d)
    with open(out + ".json", "w") as f:
        json.dump(res, f, indent=1)
    with open(out + ".txt", "w") as f:
        for d in res:
            for k, v in d.items():
                f.write("%-90s %s\n" % (k, v))
            f.write("\n")
    print("wrote %s.json / .txt (%d launches)" % (out, len(res)))


if __name__ == "__main__":
    main()

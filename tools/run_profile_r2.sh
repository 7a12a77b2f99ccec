#!/usr/bin/env bash
# tools/run_profile_r2.sh -- on the GPU box: timings without ncu, then ONE ncu --set full capture of every kernel
# flavour (tools/profile_r2.py), summarised on the spot; then the launch list of the bench command.
set -x
mkdir -p gpurun_out
python tools/profile_r2.py > gpurun_out/r2_kernels_timing.txt 2>&1 || { tail -20 gpurun_out/r2_kernels_timing.txt; exit 1; }
cat gpurun_out/r2_kernels_timing.txt
CCGPU_MANIFEST=gpurun_out/r2_manifest.json ncu --set full --clock-control none -f -o gpurun_out/r2_kernels \
    python tools/profile_r2.py > gpurun_out/r2_ncu.log 2>&1
tail -3 gpurun_out/r2_ncu.log
python tools/ncu_summary.py gpurun_out/r2_kernels.ncu-rep gpurun_out/r2_kernels --manifest gpurun_out/r2_manifest.json
ls -la gpurun_out/r2_kernels.ncu-rep
# keep the report only if it fits the transfer limit
[ $(stat -c %s gpurun_out/r2_kernels.ncu-rep) -gt 40000000 ] && rm -f gpurun_out/r2_kernels.ncu-rep
python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err || tail -5 gpurun_out/r2_bench_n1.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/r2_ncu_bench.log 2>&1
tail -2 gpurun_out/r2_ncu_bench.log

#!/bin/bash
# Round evidence run (on the GPU box): bench lines, kernel micro-benchmarks, ncu launch list and full captures.
# Everything lands in gpurun_out/; tools/ncu_summary.py condenses the ncu exports into profiles/.
#   usage: tools/round_profile.sh [tag]     (tag defaults to r02)
set -u
O=gpurun_out
T=${1:-r02}
python bench.py --impl reference --steps 5 --warmup 3 > $O/${T}_bench_ref_cfg2.json 2> $O/${T}_bench_ref_cfg2.err
python bench.py > $O/${T}_bench_cfg2.json 2> $O/${T}_bench_cfg2.err; tail -c 300 $O/${T}_bench_cfg2.err
python tools/chainbench.py > $O/${T}_chainbench.log 2>&1
python tools/chainbench.py cfg2 > $O/${T}_chainbench_cfg2.log 2>&1
python tools/kbench.py all > $O/${T}_kbench.log 2>&1
# ncu: launch list of the bench command (after the same command exited 0 without ncu)
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-workloads > $O/${T}_plain_bench.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/${T}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-workloads > $O/${T}_ncu_launch.log 2>&1
# ncu: full captures of the hot kernels (cfg2 chains, one 1024^2 chain, SSIM)
python tools/one_dwt.py 64 304 304 db3 symmetric 3 4 > $O/${T}_plain_a.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k "regex:tma_kernel|owner_kernel" -s 4 -c 2 -f -o $O/${T}_prof_cfg2_chain \
    python tools/one_dwt.py 64 304 304 db3 symmetric 3 4 > $O/${T}_ncu_a.log 2>&1
python tools/one_dwt.py 64 1024 1024 db3 symmetric 3 3 > $O/${T}_plain_b.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:stream_kernel -s 2 -c 2 -f -o $O/${T}_prof_1024_chain \
    python tools/one_dwt.py 64 1024 1024 db3 symmetric 3 3 > $O/${T}_ncu_b.log 2>&1
python tools/prof_run.py ssim 2 > $O/${T}_plain_d.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:ssim -s 2 -c 3 -f -o $O/${T}_prof_ssim \
    python tools/prof_run.py ssim 2 > $O/${T}_ncu_d.log 2>&1
ls -la $O | tail -30

#!/bin/bash
# Round evidence run (on the GPU box): tests, bench lines, kernel micro-benchmarks, ncu launch list and full captures.
# Everything lands in gpurun_out/; tools/ncu_summary.py condenses the ncu exports into profiles/.
set -u
O=gpurun_out
python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; tail -3 $O/pytest_gpu.log
python bench.py --impl reference --steps 5 --warmup 3 > $O/bench_ref_cfg2.json 2> $O/bench_ref_cfg2.err
python bench.py > $O/bench_cfg2.json 2> $O/bench_cfg2.err; tail -c 300 $O/bench_cfg2.err
python bench.py --workload cfg1 --steps 100 > $O/bench_cfg1.json 2> $O/bench_cfg1.err; tail -c 300 $O/bench_cfg1.err
python bench.py --workload cfg3 --steps 50 > $O/bench_cfg3.json 2> $O/bench_cfg3.err; tail -c 300 $O/bench_cfg3.err
python tools/chainbench.py > $O/chainbench.log 2>&1
python tools/kbench.py all > $O/kbench.log 2>&1
python tools/pcie_probe.py > $O/pcie.log 2>&1
# ncu: launch list of the bench command (after the same command exited 0 above without ncu)
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/plain_bench.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/ncu_launch.log 2>&1
# ncu: full captures of the hot kernels
python tools/prof_dwt.py 64 304 db3 symmetric 3 > $O/plain_a.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k "regex:stream_kernel|owner_kernel" -s 2 -c 2 -f -o $O/prof_cfg2_chain \
    python tools/prof_dwt.py 64 304 db3 symmetric 3 > $O/ncu_a.log 2>&1
python tools/prof_dwt.py 64 304 db3 symmetric 1 > $O/plain_b.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:stream_kernel -s 2 -c 2 -f -o $O/prof_cfg2_level1 \
    python tools/prof_dwt.py 64 304 db3 symmetric 1 > $O/ncu_b.log 2>&1
python tools/prof_dwt.py 64 1024 db3 symmetric 1 > $O/plain_c.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:stream_kernel -s 2 -c 2 -f -o $O/prof_1024_level1 \
    python tools/prof_dwt.py 64 1024 db3 symmetric 1 > $O/ncu_c.log 2>&1
python tools/prof_run.py ssim 2 > $O/plain_d.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:ssim_stream -s 2 -c 2 -f -o $O/prof_ssim \
    python tools/prof_run.py ssim 2 > $O/ncu_d.log 2>&1
ls -la $O | tail -30

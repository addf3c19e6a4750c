#!/bin/bash
# Evidence round for the whole-program drop-in (config 1) and the overlap aligner (f-1); run under gpurun.
mkdir -p gpurun_out
(timeout 150 python -m pytest tests/test_overlap_align.py -m gpu -x -q 2>&1 | tail -5) > gpurun_out/r3_align_tests.log
timeout 200 python tools/align_bench.py --dir oracle/_ref/example_300 --out gpurun_out/align_300.json > gpurun_out/align_300.log 2>&1
timeout 300 python tools/example_run.py --dir oracle/_ref/example_300 --gpu-align --out gpurun_out/example_300.json > gpurun_out/example_300.log 2>&1
timeout 150 ncu --set full --clock-control none --import-source on -k regex:ovl_kernel -c 1 -f -o gpurun_out/prof_ovl \
    python tools/align_bench.py --dir oracle/_ref/example_300 --n 1184 --repeat 1 --cpu-sample 0 > gpurun_out/ncu_ovl.log 2>&1
if [ -s oracle/_ref/example/corrected.ref.fa ]; then
  timeout 400 python tools/example_run.py --dir oracle/_ref/example --gpu-align --skip-ref --out gpurun_out/example_full.json > gpurun_out/example_full.log 2>&1
fi
tail -3 gpurun_out/r3_align_tests.log; tail -2 gpurun_out/align_300.log; tail -30 gpurun_out/example_300.log | head -60; tail -3 gpurun_out/ncu_ovl.log; tail -25 gpurun_out/example_full.log

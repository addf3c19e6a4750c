#!/bin/bash
# last GPU session of round 2: binding tests on the final binary, smoke, launch list + one full capture per changed kernel
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_example_binary.py tests/test_overlap_align.py -m gpu -x -q > gpurun_out/pytest_binary.log 2>&1; echo "binary tests rc=$?"; tail -3 gpurun_out/pytest_binary.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -5 gpurun_out/smoke.log
export VGC_GROUPS=1
BENCH="python bench.py --steps 1 --warmup 1 --targets 100 --no-cpu-baseline"
timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -c 4000 --csv \
  --log-file gpurun_out/r02_launches.csv $BENCH > gpurun_out/r02_ncu_launches.log 2>&1
python tools/launch_summary.py gpurun_out/r02_launches.csv | tee gpurun_out/r02_launch_summary.txt
timeout 300 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:update_kernel" -s 20 -c 1 -f \
  -o gpurun_out/r02_prof_update $BENCH > gpurun_out/r02_ncuk_update.log 2>&1
ls -la gpurun_out/r02_prof_update.ncu-rep

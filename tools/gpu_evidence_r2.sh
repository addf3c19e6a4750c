#!/bin/bash
# Round-2 evidence (run under gpurun): launch list, DRAM traffic of the align kernels over a whole (small) pass, and
# one `ncu --set full` capture per kernel, all on the final binary.  Numbers printed under ncu are never bench values.
mkdir -p gpurun_out
export VGC_GROUPS=1
BENCH="python bench.py --steps 1 --warmup 1 --targets 100 --no-cpu-baseline"
# 1. launch list (cold-cache, serialised): share of each kernel in a pass
timeout 1200 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -c 4000 --csv \
  --log-file gpurun_out/r02_launches.csv $BENCH > gpurun_out/r02_ncu_launches.log 2>&1
python tools/launch_summary.py gpurun_out/r02_launches.csv | tee gpurun_out/r02_launch_summary.txt
# 2. DRAM bytes of every align_kernel launch of one pass (the second pass of the run: skip the warm-up's launches)
timeout 1200 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
  -k regex:align -c 4000 --csv --log-file gpurun_out/r02_align_dram.csv $BENCH > gpurun_out/r02_ncu_dram.log 2>&1
grep -o '"dp_cells_per_launch": [0-9]*' gpurun_out/r02_ncu_dram.log | tail -1
# 3. full captures
cap() {  # name, demangled regex, skip
  timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$2" -s $3 -c 1 -f \
    -o gpurun_out/r02_prof_$1 $BENCH > gpurun_out/r02_ncuk_$1.log 2>&1
  ls -la gpurun_out/r02_prof_$1.ncu-rep
}
cap align8 'align_kernel<.int.8,..bool.0>' 25
cap align8_round 'align_kernel<.int.8,..bool.0>' 38
cap align10 'align_kernel<.int.10,..bool.0>' 20
cap align8_sw 'align_kernel<.int.8,..bool.1>' 3
cap sort 'sort_kernel' 32
cap update 'update_kernel' 20

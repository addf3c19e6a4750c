#!/bin/bash
# One GPU session: parity tests, smoke, bench (both arms), ncu launch list (+ optional full capture).
# Usage (under gpurun): bash tools/gpu_round.sh [tests] [bench] [ncu] [ncufull]
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
for what in "$@"; do
  case $what in
    tests)
      timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
      tail -5 gpurun_out/pytest_gpu.log
      timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
      tail -6 gpurun_out/smoke.log ;;
    bench)
      timeout 900 python bench.py --steps ${STEPS:-5} --warmup ${WARMUP:-3} ${BENCH_ARGS} > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
      tail -2 gpurun_out/bench.err; cat gpurun_out/bench.json ;;
    benchref)
      timeout 600 python bench.py --impl reference --steps ${STEPS:-5} --warmup ${WARMUP:-3} > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "benchref rc=$?"
      cat gpurun_out/bench_ref.json ;;
    ncu)
      VGC_GROUPS=${NCU_GROUPS:-1} timeout 1500 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -s ${NCU_SKIP:-0} -c ${NCU_COUNT:-3000} --csv --log-file gpurun_out/launches.csv \
        python bench.py --steps 1 --warmup 1 --targets ${NCU_TARGETS:-100} --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; echo "ncu rc=$?"
      python tools/launch_summary.py gpurun_out/launches.csv | tee gpurun_out/launch_summary.txt ;;
    ncugraph)
      timeout 1500 ncu --set full --clock-control none --import-source on -k regex:graph_kernel -s ${NCU_SKIP:-100} -c 1 -f -o gpurun_out/prof_graph \
        python bench.py --steps 1 --warmup 1 --targets ${NCU_TARGETS:-100} --no-cpu-baseline > gpurun_out/ncugraph.log 2>&1; echo "ncugraph rc=$?"
      ls -la gpurun_out/prof_graph.ncu-rep ;;
    ncuk)
      # generic: NCU_KERNEL regex, NCU_SKIP, output gpurun_out/prof_${NCU_NAME}
      VGC_GROUPS=${NCU_GROUPS:-1} timeout 1500 ncu --set full --clock-control none --import-source on -k regex:${NCU_KERNEL} -s ${NCU_SKIP:-100} -c 1 -f -o gpurun_out/prof_${NCU_NAME:-k} \
        python bench.py --steps 1 --warmup 1 --targets ${NCU_TARGETS:-100} --no-cpu-baseline > gpurun_out/ncuk.log 2>&1; echo "ncuk rc=$?"
      ls -la gpurun_out/prof_${NCU_NAME:-k}.ncu-rep ;;
    ncufill)
      timeout 1500 ncu --set full --clock-control none --import-source on -k regex:fill_kernel -s ${NCU_SKIP:-100} -c 1 -f -o gpurun_out/prof_fill \
        python bench.py --steps 1 --warmup 1 --targets ${NCU_TARGETS:-100} --no-cpu-baseline > gpurun_out/ncufill.log 2>&1; echo "ncufill rc=$?"
      ls -la gpurun_out/prof_fill.ncu-rep ;;
    ncufull)
      timeout 1500 ncu --set full --clock-control none --import-source on -k regex:poa_window -s 1 -c 1 -f -o gpurun_out/prof \
        python bench.py --steps 1 --warmup 1 --targets ${NCU_TARGETS:-100} --no-cpu-baseline > gpurun_out/ncufull.log 2>&1; echo "ncufull rc=$?"
      ls -la gpurun_out/prof.ncu-rep ;;
  esac
done

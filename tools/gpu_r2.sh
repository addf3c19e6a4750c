#!/bin/bash
# Round-2 GPU session helper (run under gpurun): bash tools/gpu_r2.sh [parity] [tests] [smoke] [bench] [ncu] ...
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
for what in "$@"; do
  case $what in
    parity)
      timeout ${T_PARITY:-900} python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/pytest_parity.log 2>&1; echo "parity rc=$?" >> gpurun_out/pytest_parity.log
      tail -15 gpurun_out/pytest_parity.log ;;
    tests)
      timeout ${T_TESTS:-1500} python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
      tail -15 gpurun_out/pytest_gpu.log ;;
    smoke)
      timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
      tail -6 gpurun_out/smoke.log ;;
    bench)
      timeout 900 python bench.py --steps ${STEPS:-5} --warmup ${WARMUP:-3} ${BENCH_ARGS} > gpurun_out/bench${TAG}.json 2> gpurun_out/bench${TAG}.err; echo "bench rc=$?"
      tail -3 gpurun_out/bench${TAG}.err; cat gpurun_out/bench${TAG}.json ;;
    benchsmall)
      timeout 600 python bench.py --steps 3 --warmup 3 --targets 400 --no-cpu-baseline > gpurun_out/bench_small${TAG}.json 2> gpurun_out/bench_small${TAG}.err; echo "benchsmall rc=$?"
      tail -3 gpurun_out/bench_small${TAG}.err; cat gpurun_out/bench_small${TAG}.json ;;
    benchref)
      timeout 600 python bench.py --impl reference --steps ${STEPS:-5} --warmup ${WARMUP:-3} > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "benchref rc=$?"
      cat gpurun_out/bench_ref.json ;;
    ncu)
      VGC_GROUPS=${NCU_GROUPS:-1} timeout 1500 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -s ${NCU_SKIP:-0} -c ${NCU_COUNT:-3000} --csv --log-file gpurun_out/launches.csv \
        python bench.py --steps 1 --warmup 1 --targets ${NCU_TARGETS:-100} --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; echo "ncu rc=$?"
      python tools/launch_summary.py gpurun_out/launches.csv | tee gpurun_out/launch_summary.txt ;;
    ncuk)
      # generic: NCU_KERNEL regex, NCU_SKIP, output gpurun_out/prof_${NCU_NAME}
      VGC_GROUPS=${NCU_GROUPS:-1} timeout 1500 ncu --set full --clock-control none --import-source on --kernel-name-base ${NCU_BASE:-function} -k regex:${NCU_KERNEL} -s ${NCU_SKIP:-100} -c 1 -f -o gpurun_out/prof_${NCU_NAME:-k} \
        python bench.py --steps 1 --warmup 1 --targets ${NCU_TARGETS:-100} --no-cpu-baseline > gpurun_out/ncuk_${NCU_NAME:-k}.log 2>&1; echo "ncuk rc=$?"
      ls -la gpurun_out/prof_${NCU_NAME:-k}.ncu-rep ;;
  esac
done

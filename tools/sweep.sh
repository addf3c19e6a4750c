#!/bin/bash
# usage: bash tools/sweep.sh VAR v1 v2 ... : run the bench (2 steps) once per value of env VAR
var=$1; shift
mkdir -p gpurun_out
for v in "$@"; do
  env $var=$v timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline ${BENCH_ARGS} 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$var=$v', 'value %.0f e2e %.0f frac %.3f launches %d' % (d['value'], d['e2e']['value'], d['roofline']['frac'], d['gpu_launches']), d['phase_share'])
" | tee -a gpurun_out/sweep.log
done

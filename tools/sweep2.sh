#!/bin/bash
# usage: bash tools/sweep2.sh "ENV=val ENV2=val2 -- bench args" ... : one short bench per configuration
mkdir -p gpurun_out
for cfg in "$@"; do
  envs="${cfg%%--*}"; args="${cfg#*--}"
  env $envs timeout 900 python bench.py --steps 2 --warmup 1 --no-cpu-baseline $args 2>gpurun_out/sweep.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('''$cfg''', '| value %.0f e2e %.0f frac %.3f launches %d ms/step %.0f' % (d['value'], d['e2e']['value'], d['roofline']['frac'], d['gpu_launches'], d['ms_per_step']), d['phase_share'], 'RAW', d.get('phase_raw'), 'aln', d.get('alignments'), 'relaunched', d.get('relaunched_windows'), 'cells', d['roofline']['dp_cells_per_launch'])
" | tee -a gpurun_out/sweep.log
done

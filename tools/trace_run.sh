# timing trace of the e2e pipeline (VGC_VERBOSE + BENCH_TRACE): all cores, then 4 cores as at N = 8 on a 32-core host
mkdir -p gpurun_out
one() {  # tag, env...
  local tag=$1; shift
  env VGC_VERBOSE=1 BENCH_TRACE=1 "$@" python bench.py --steps ${STEPS:-3} --warmup 3 --no-cpu-baseline > gpurun_out/tr_$tag.json 2> gpurun_out/tr_$tag.err
  grep -E "^\[vgc\] (collect|pass setup|after)|^\[bench\]" gpurun_out/tr_$tag.err | tail -${TAIL:-14}
  python -c "
import json; d=json.load(open('gpurun_out/tr_$tag.json')); print('$tag value', d['value'], 'e2e', d['e2e']['value'], 'e2e ms', d['e2e']['ms_per_step'], 'resident ms', d['ms_per_step'], 'kernel ms', d['roofline']['kernel_ms_per_launch'])"
}
echo "=== N=1, all cores"; one full
if [ -n "$FOUR" ]; then echo "=== N=1, 4 cores"; one 4c VGC_HOST_THREADS=4 taskset -c 0-3; fi

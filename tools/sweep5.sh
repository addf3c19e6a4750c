#!/bin/bash
# A/B of library variants (run under gpurun): build variants of libvgc.so with -D switches into exp_libs/libvgc_<tag>.so
# (nvcc line of vechat_b200/build.py + the switch), then `tools/sweep5.sh [parity:<tag>] default <tag> default <tag> ...`:
# "parity:<tag>" runs the GPU parity tests against that variant, every other word is one short bench run (VGC_LIB).
# Same box, alternating order: box-to-box variation is ~1 %, run-to-run on one box ~0.1 %.
mkdir -p gpurun_out
run() {
  local tag="$1"; shift
  env "$@" timeout 600 python bench.py --steps ${STEPS:-3} --warmup 3 --no-cpu-baseline > gpurun_out/sw_$tag.json 2> gpurun_out/sw_$tag.err
  python - "$tag" <<'PY'
import json,sys
tag=sys.argv[1]
try:
    d=json.load(open("gpurun_out/sw_%s.json"%tag))
    pr=d["phase_raw"]
    print("%-12s value %8.0f e2e %8.0f kernel_ms %7.1f frac %.4f fill %6.0fG trace %6.0fG addaln %5.0fG topo %5.0fG rowprog %5.0fG"%(tag,d["value"],d["e2e"]["value"],d["roofline"]["kernel_ms_per_launch"],d["roofline"]["frac"],pr["fill"]/1e9,pr["traceback"]/1e9,pr["add_alignment"]/1e9,pr["toposort"]/1e9,pr["rowprog"]/1e9))
except Exception as e:
    print(tag,"FAILED",e)
PY
}
for v in "$@"; do
  case $v in
    parity:*) t=${v#parity:}; VGC_LIB=$PWD/exp_libs/libvgc_$t.so timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3 ;;
    default) run default ;;
    *) run "$v" VGC_LIB=$PWD/exp_libs/libvgc_$v.so ;;
  esac
done

#!/bin/bash
mkdir -p gpurun_out
run() {
  local tag="$1"; shift
  env "$@" timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/sw_$tag.json 2> gpurun_out/sw_$tag.err
  python - "$tag" <<'PY'
import json,sys
tag=sys.argv[1]
try:
    d=json.load(open("gpurun_out/sw_%s.json"%tag))
    pr=d["phase_raw"]
    print("%-24s value %8.0f e2e %8.0f kernel_ms %7.1f fill %6.0fG trace %6.0fG addaln %5.0fG parity %s"%(tag,d["value"],d["e2e"]["value"],d["roofline"]["kernel_ms_per_launch"],pr["fill"]/1e9,pr["traceback"]/1e9,pr["add_alignment"]/1e9,d.get("parity",{}).get("mismatches")))
except Exception as e:
    print(tag,"FAILED",e)
PY
}
for v in "$@"; do
  case $v in
    base) run base VGC_UPDATE_SMEM=6144 ;;
    c20) run c20 VGC_UPDATE_SMEM=6144 VGC_LIB=$PWD/exp_libs/libvgc_20_4.so ;;
    c21) run c21 VGC_UPDATE_SMEM=6144 VGC_LIB=$PWD/exp_libs/libvgc_21_4.so ;;
    c24) run c24 VGC_UPDATE_SMEM=6144 VGC_LIB=$PWD/exp_libs/libvgc_24_2.so ;;
    *) run "$v" VGC_UPDATE_SMEM=6144 VGC_LIB=$PWD/exp_libs/libvgc_$v.so ;;
  esac
done

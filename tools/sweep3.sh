#!/bin/bash
# experiment sweep (run under gpurun): env variants of the same bench, value + kernel ms only
mkdir -p gpurun_out
run() {
  local tag="$1"; shift
  env "$@" timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/sw_$tag.json 2> gpurun_out/sw_$tag.err
  python - "$tag" <<'PY'
import json,sys
tag=sys.argv[1]
try:
    d=json.load(open("gpurun_out/sw_%s.json"%tag))
    print("%-28s value %8.0f  e2e %8.0f  kernel_ms %7.1f launch_ms %5.1f"%(tag,d["value"],d["e2e"]["value"],d["roofline"]["kernel_ms_per_launch"],d["phase_raw"]["host_launch_ms"]))
except Exception as e:
    print(tag,"FAILED",e)
PY
}
run base A=1
run conn32 CUDA_DEVICE_MAX_CONNECTIONS=32
run conn32_g64 CUDA_DEVICE_MAX_CONNECTIONS=32 VGC_GROUPS=64
run conn32_g24 CUDA_DEVICE_MAX_CONNECTIONS=32 VGC_GROUPS=24
run conn16 CUDA_DEVICE_MAX_CONNECTIONS=16
run conn32_usm6k CUDA_DEVICE_MAX_CONNECTIONS=32 VGC_UPDATE_SMEM=6144

# does the NCCL gather of the corrected reads (e2e leg, N > 1) steal SMs from the pass?  e2e kernel_ms vs resident, with
# NCCL's default channel count and with NCCL_MAX_NCHANNELS=2
mkdir -p gpurun_out
for v in default ch2; do
  if [ $v = ch2 ]; then export NCCL_MAX_NCHANNELS=2; fi
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline --no-hap2-leg > gpurun_out/n2_$v.json 2> gpurun_out/n2_$v.err
  python - <<PY
import json
for l in open("gpurun_out/n2_$v.json"):
    if l.startswith("{"):
        d=json.loads(l); e=d["e2e"]
        print("$v", "value", round(d["value"]), "e2e", round(e["value"]), "e2e ms", round(e["ms_per_step"],1), "resident ms", round(d["ms_per_step"],1), "e2e kernel_ms", round(e["kernel_ms"],1), "resident kernel ms", round(d["roofline"]["kernel_ms_per_launch"],1))
PY
done

#!/bin/bash
# N-GPU sweep of the fused all-gather schedule (run under `gpurun --gpus N`): one bench line per (sub-batches, gather SMs).
# usage: bash tools/scale_sweep.sh N tag "sub:sms sub:sms ..." [extra bench args]
N=$1; tag=$2; combos=$3; shift 3
mkdir -p gpurun_out
port=29540
for c in $combos; do
  sub=${c%%:*}; sms=${c##*:}
  port=$((port+1))
  mode=fused; ctas=24; gsms=$sms
  if [ "$sms" = "ce" ]; then mode=ce; gsms=0; fi
  case "$sms" in push*) mode=push; ctas=${sms#push}; gsms=0;; esac      # "push24" = push variant with 24 CTAs
  MAMIMO_GATHER_MODE=$mode MAMIMO_PUSH_CTAS=${ctas:-24} MAMIMO_GATHER_SUB=$sub MAMIMO_GATHER_SMS=$gsms python -m torch.distributed.run --nnodes=1 --nproc-per-node $N \
      --master-addr 127.0.0.1 --master-port $port bench.py --gpus $N --steps 20 --warmup 3 --no-cpu-baseline --no-e2e "$@" \
      > gpurun_out/${tag}_n${N}_sub${sub}_sms${sms}.json 2> gpurun_out/${tag}_n${N}_sub${sub}_sms${sms}.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/${tag}_n${N}_sub${sub}_sms${sms}.json").read().strip().splitlines()[-1])
    print("N=$N sub=$sub sms=$sms value %.0f pkt/s  %.3f ms/step  compute-only %.0f  fused==nccl %s  parity %s" % (
        d["value"], d["ms_per_step"], d.get("value_compute_only", 0), d.get("fused_gather_equals_nccl"), (d.get("parity") or {}).get("max")))
except Exception as ex:
    print("N=$N sub=$sub sms=$sms FAILED", ex)
    print(open("gpurun_out/${tag}_n${N}_sub${sub}_sms${sms}.err").read()[-1500:])
PY
done

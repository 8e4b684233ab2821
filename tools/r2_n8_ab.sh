#!/bin/bash
# 8-GPU A/B of the multi-GPU switches (each bench ~25 s; whole script ~3 min of box time = ~24 GPU-minutes at --gpus 8):
#   gpurun --gpus 8 --timeout 400 -- bash tools/r2_n8_ab.sh
# columns: B200_GEMM_SCHED (CLC tile scheduler)  B200_WGRAD_STREAM  B200_PEER_COMM (0 NCCL | ag NCCL RS + copy-engine AG)
mkdir -p gpurun_out
port=29600
for cfg in "0 0 0" "1 0 0" "1 1 0" "1 0 ag"; do
  set -- $cfg
  port=$((port+1))
  B200_GEMM_SCHED=$1 B200_WGRAD_STREAM=$2 B200_PEER_COMM=$3 B200_BENCH_WATCHDOG_S=90 timeout 100 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 \
    --master-addr 127.0.0.1 --master-port $port bench.py --gpus 8 --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/r2_n8_$1_$2_$3.json 2> gpurun_out/r2_n8_$1_$2_$3.err
  python - "$1" "$2" "$3" <<'PY'
import json, sys
try:
    d = json.load(open(f"gpurun_out/r2_n8_{sys.argv[1]}_{sys.argv[2]}_{sys.argv[3]}.json"))
    print(f"sched={sys.argv[1]} wgrad={sys.argv[2]} peer={sys.argv[3]}: {d['ms_per_step']:.2f} ms {d['value']:.0f} tok/s  e2e {d['e2e']['value']:.0f}  clocks {d['clocks']['sm_mhz']}")
except Exception as e:
    print("FAILED", sys.argv[1:], e)
PY
done

#!/bin/bash
# A/B of the two stream-overlap switches on one GPU: B200_WGRAD_STREAM (weight-gradient GEMMs on their own stream) and
# B200_SIDE_BLOCKS (occupancy cap of the side-stream AdamW / grad-norm sweeps).  Identical final_loss / final_grad_norm across
# the runs = no ordering hazard; ms_per_step is the result.   usage: gpurun -- bash tools/overlap_sweep.sh
mkdir -p gpurun_out
B200_WGRAD_STREAM=1 B200_SIDE_BLOCKS=2 timeout 400 python -m pytest tests/test_engine_gpu.py -x -q -m gpu > gpurun_out/ov_tests.log 2>&1
echo "tests rc=$?"; tail -2 gpurun_out/ov_tests.log
for cfg in "0 0" "0 2" "1 0" "1 2"; do
  set -- $cfg
  B200_WGRAD_STREAM=$1 B200_SIDE_BLOCKS=$2 timeout 200 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/ov_$1_$2.json 2> gpurun_out/ov_$1_$2.err
  python - "$1" "$2" <<'PY'
import json, sys
try:
    d = json.load(open(f"gpurun_out/ov_{sys.argv[1]}_{sys.argv[2]}.json"))
    print(f"wgrad_stream={sys.argv[1]} side_blocks={sys.argv[2]}: {d['ms_per_step']:.2f} ms  {d['value']:.0f} tok/s  e2e {d['e2e']['value']:.0f}  gemm {d['roofline']['achieved']:.0f} TF  loss {d['final_loss']!r} gnorm {d.get('final_grad_norm')!r} clocks {d['clocks']['sm_mhz']}")
except Exception as e:
    print("FAILED", sys.argv[1:], e)
PY
done

#!/bin/bash
# 2-GPU check of the weight-gradient stream next to the NCCL collectives: world-2 parity test with it on, then bench off/on.
mkdir -p gpurun_out
B200_WGRAD_STREAM=1 timeout 60 python -m pytest tests/test_dist_gpu.py -q -m gpu -k "nccl" > gpurun_out/n2wg_tests.log 2>&1; echo "dist tests rc=$?"; tail -2 gpurun_out/n2wg_tests.log
for wg in 0 1; do
  B200_WGRAD_STREAM=$wg timeout 45 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((29511+wg)) bench.py --gpus 2 --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/n2wg_$wg.json 2> gpurun_out/n2wg_$wg.err; echo "bench wg=$wg rc=$?"
  python -c "
import json
d=json.load(open('gpurun_out/n2wg_$wg.json')); print('wg=$wg', round(d['ms_per_step'],2),'ms', round(d['value']),'tok/s e2e', round(d['e2e']['value']), 'clk', d['clocks']['sm_mhz'])"
done

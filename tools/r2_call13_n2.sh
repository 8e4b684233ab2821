#!/bin/bash
# two GPUs: the reference's DTensor / FSDP2 path (world 2) vs ours under the same unmodified recipe, Llama-3-8B config, 100 steps
mkdir -p gpurun_out
O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
OMP_NUM_THREADS=16 timeout 170 $TR --master-port 29801 tools/ref_gpu_run.py --strategy fsdp2 --config 8b --steps 60 --out $O/ref_8b_n2.json > $O/ref_8b_n2.log 2>&1; echo "ref 8b n2 rc=$?"; tail -1 $O/ref_8b_n2.log | cut -c1-300
timeout 100 $TR --master-port 29802 tools/ref_gpu_run.py --strategy b200_sharded --config 8b --steps 60 --out $O/b200_8b_n2.json > $O/b200_8b_n2.log 2>&1; echo "b200 8b n2 rc=$?"; tail -1 $O/b200_8b_n2.log | cut -c1-300
python tools/ref_gpu_run.py --compare $O/ref_8b_n2.json $O/b200_8b_n2.json --md $O/r2_parity_n2.md | sed -n 1,10p
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/ref_8b_n2.json")); r = d.get("ref_inv_freq") or {}
    print("ref inv_freq:", d.get("ref_inv_freq_was_initialised"), d.get("ref_inv_freq_max_rel_dev_from_formula"), r.get("dtype"), r.get("device"), "bf16-rounded formula:", r.get("stock_equals_bf16_rounded_formula"), (r.get("stock") or [])[:4], (r.get("formula") or [])[:4])
except Exception as e:
    print("no ref record", e)
PY

#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -q -m gpu -x -k "attention" > $O/r2_attn_tests2.log 2>&1; echo "attention tests rc=$?"; tail -3 $O/r2_attn_tests2.log
for v in 2 3 4 2 3 4; do timeout 120 python tools/attn_vs_cudnn.py --variant $v --ours-only 2>&1 | tail -1; done
timeout 400 python tools/ref_gpu_run.py --strategy fsdp2 --config 8b --steps 100 --sdpa-backend flash --out $O/ref_8b_n1_flash.json > $O/ref_8b_n1_flash.log 2>&1; echo "ref (flash sdpa) rc=$?"; tail -1 $O/ref_8b_n1_flash.log
timeout 400 python tools/ref_gpu_run.py --strategy fsdp2 --config 8b --steps 100 --sdpa-backend cudnn --out $O/ref_8b_n1_cudnn.json > $O/ref_8b_n1_cudnn.log 2>&1; echo "ref (cudnn sdpa) rc=$?"; tail -1 $O/ref_8b_n1_cudnn.log
python tools/ref_gpu_run.py --compare $O/ref_8b_n1_cudnn.json $O/ref_8b_n1_flash.json | sed -n 1,9p

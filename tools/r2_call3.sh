#!/bin/bash
# Round-2 GPU call 3 (one GPU): new attention forward (P in TMEM), cuDNN comparison, ncu captures, bf16 noise floor of the 8B forward,
# the reference's own kernel-choice noise over 100 training steps.
mkdir -p gpurun_out
O=gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -q -m gpu -x -k "attention" > $O/r2_attn_tests.log 2>&1; echo "attention tests rc=$?"; tail -3 $O/r2_attn_tests.log
for v in 1 2; do timeout 120 python tools/attn_vs_cudnn.py --variant $v --ours-only 2>&1 | tail -1; done
timeout 200 python tools/attn_vs_cudnn.py --variant 2 --md $O/r2_attn_vs_cudnn.md 2>&1 | tail -8
for v in 1 2; do
  timeout 200 ncu --set full --clock-control none --import-source on -k regex:attn_fwd -c 1 -o $O/r2_attn_fwd_v$v -f python tools/attn_vs_cudnn.py --variant $v --ours-only --iters 1 > $O/r2_ncu_fwd_v$v.log 2>&1; echo "ncu fwd v$v rc=$?"
done
timeout 200 ncu --set full --clock-control none --import-source on -k regex:attn_bwd_tc -c 1 -o $O/r2_attn_bwd -f python tools/attn_vs_cudnn.py --ours-only --iters 1 > $O/r2_ncu_bwd.log 2>&1; echo "ncu bwd rc=$?"

timeout 300 python tools/ref_gpu_run.py --strategy fsdp2 --config 8b --dtype float32 --fwd-only 8 --steps 1 --out $O/noise_ref_fp32.json > $O/noise_ref_fp32.log 2>&1; echo "fp32 ref fwd rc=$?"; tail -1 $O/noise_ref_fp32.log
timeout 300 python tools/ref_gpu_run.py --strategy fsdp2 --config 8b --fwd-only 8 --steps 1 --out $O/noise_ref_bf16.json > $O/noise_ref_bf16.log 2>&1; echo "bf16 ref fwd rc=$?"; tail -1 $O/noise_ref_bf16.log
timeout 300 python tools/ref_gpu_run.py --strategy b200_sharded --config 8b --fwd-only 8 --steps 1 --out $O/noise_ours.json > $O/noise_ours.log 2>&1; echo "ours fwd rc=$?"; tail -1 $O/noise_ours.log
python tools/ref_gpu_run.py --noise $O/noise_ref_fp32.json $O/noise_ref_bf16.json $O/noise_ours.json --md $O/r2_noise_floor.md | tail -16
timeout 400 python tools/ref_gpu_run.py --strategy fsdp2 --config 8b --steps 100 --sdpa-backend efficient --out $O/ref_8b_n1_efficient.json > $O/ref_8b_n1_efficient.log 2>&1; echo "ref (mem-efficient sdpa) rc=$?"; tail -1 $O/ref_8b_n1_efficient.log
timeout 400 python tools/ref_gpu_run.py --strategy fsdp2 --config 8b --steps 100 --out $O/ref_8b_n1_b.json > $O/ref_8b_n1_b.log 2>&1; echo "ref (default, second run) rc=$?"; tail -1 $O/ref_8b_n1_b.log
python tools/ref_gpu_run.py --compare $O/ref_8b_n1_b.json $O/ref_8b_n1_efficient.json | sed -n 1,9p

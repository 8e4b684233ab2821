#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
timeout 200 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "attention" > $O/r2_attn_tests4.log 2>&1; echo "attention tests rc=$?"; grep -E "^FAILED|passed|failed" $O/r2_attn_tests4.log | tail -5 | cut -c1-200
timeout 200 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "bit_reproducible" > $O/r2_attn_tests5.log 2>&1; echo "repro again rc=$?"; tail -1 $O/r2_attn_tests5.log
timeout 100 python tools/attn_vs_cudnn.py --variant 2 --ours-only 2>&1 | tail -1

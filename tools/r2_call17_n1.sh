#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests -q -m gpu > gpurun_out/r2_final_gpu_suite5.log 2>&1; echo "gpu suite rc=$?"; grep -E "^FAILED|passed|failed" gpurun_out/r2_final_gpu_suite5.log | tail -5 | cut -c1-250

#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
timeout 200 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "attention" > $O/r2_attn_tests3.log 2>&1; echo "attention tests rc=$?"; tail -3 $O/r2_attn_tests3.log | cut -c1-300
timeout 600 python -m pytest tests -q -m gpu > $O/r2_final_gpu_suite2.log 2>&1; echo "gpu suite rc=$?"; grep -E "^FAILED|passed|failed" $O/r2_final_gpu_suite2.log | tail -8 | cut -c1-300
for v in 2 1; do timeout 100 python tools/attn_vs_cudnn.py --variant $v --ours-only 2>&1 | tail -1; done
timeout 300 python bench.py --steps 10 --warmup 3 > $O/r2_bench_n1b.json 2> $O/r2_bench_n1b.err; echo "bench rc=$?"
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/r2_bench_n1b.json"))
    print(f"N=1: {d['ms_per_step']:.2f} ms {d['value']:.0f} tok/s e2e {d['e2e']['value']:.0f} gemm {d['roofline']['achieved']:.0f} TF frac {d['roofline']['frac']:.3f} clocks {d['clocks']['sm_mhz']} launches {d['gpu_launches']}")
except Exception as e:
    print("FAILED", e)
PY

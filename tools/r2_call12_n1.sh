#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests -q -m gpu -rx > $O/r2_final_gpu_suite4.log 2>&1; echo "gpu suite rc=$?"; grep -E "^FAILED|^XFAIL|^XPASS|passed|failed" $O/r2_final_gpu_suite4.log | tail -8 | cut -c1-260
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 300 python bench.py --steps 10 --warmup 3 > $O/r2_bench_n1d.json 2> $O/r2_bench_n1d.err; echo "bench rc=$?"
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/r2_bench_n1d.json"))
    print(f"N=1: {d['ms_per_step']:.2f} ms {d['value']:.0f} tok/s e2e {d['e2e']['value']:.0f} gemm {d['roofline']['achieved']:.0f} TF frac {d['roofline']['frac']:.3f} clocks {d['clocks']['sm_mhz']} launches {d['gpu_launches']} cpu {d['cpu_baseline']['kind']} {d['cpu_baseline']['value']:.2f}")
except Exception as e:
    print("FAILED", e)
PY

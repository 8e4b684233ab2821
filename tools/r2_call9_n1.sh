#!/bin/bash
# one GPU: full -m gpu suite, smoke, attention A/B (4 vs 8 softmax warps), default bench, reference arm, ncu launch list of one profiled step
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests -q -m gpu > $O/r2_final_gpu_suite.log 2>&1; echo "gpu suite rc=$?"; tail -3 $O/r2_final_gpu_suite.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
for v in 2 3 2 3; do timeout 100 python tools/attn_vs_cudnn.py --variant $v --ours-only 2>&1 | tail -1; done
timeout 300 python bench.py --steps 10 --warmup 3 > $O/r2_bench_n1.json 2> $O/r2_bench_n1.err; echo "bench rc=$?"
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/r2_bench_n1.json"))
    print(f"N=1: {d['ms_per_step']:.2f} ms {d['value']:.0f} tok/s e2e {d['e2e']['value']:.0f} ({d['e2e']['api']}) gemm {d['roofline']['achieved']:.0f} TF frac {d['roofline']['frac']:.3f} clocks {d['clocks']} launches {d['gpu_launches']} cpu {d.get('cpu_baseline',{}).get('kind')} {d.get('cpu_baseline',{}).get('value')}")
except Exception as e:
    print("FAILED", e)
PY
timeout 200 python bench.py --impl reference --steps 4 --warmup 1 > $O/r2_bench_ref.json 2> $O/r2_bench_ref.err; echo "reference arm rc=$?"; cut -c1-600 $O/r2_bench_ref.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 2500 --csv --log-file $O/r2_launches.csv python bench.py --profile --steps 1 --warmup 1 --no-cpu-baseline > $O/r2_prof.log 2>&1; echo "ncu launch list rc=$?"

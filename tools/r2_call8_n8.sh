#!/bin/bash
# eight GPUs, two bench runs: default (all-gather NVLS multicast store, reduce-scatter fp32 peer loads) and reduce_dtype=bfloat16 (NVLS in-fabric reduce)
mkdir -p gpurun_out
O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
port=29750
for cfg in "float32 95 75" "bfloat16 85 65"; do
  set -- $cfg
  port=$((port+1))
  B200_COMM=nvls B200_REDUCE_DTYPE=$1 B200_BENCH_WATCHDOG_S=$3 timeout $2 $TR --master-port $port bench.py --gpus 8 --steps 10 --warmup 3 --no-cpu-baseline > $O/r2_n8b_$1.json 2> $O/r2_n8b_$1.err
  echo "bench nvls reduce=$1 rc=$?"
  python - "$1" <<'PY'
import json, sys
try:
    d = json.load(open(f"gpurun_out/r2_n8b_{sys.argv[1]}.json")); p = d["parity"]; c = p["collectives"]
    print(f"N=8 {d['config']['collectives']} reduce={sys.argv[1]}: {d['ms_per_step']:.2f} ms {d['value']:.0f} tok/s e2e {d['e2e']['value']:.0f} gemm {d['roofline']['achieved']:.0f} TF clocks {d['clocks']['sm_mhz']} launches {d['gpu_launches']} | parity ok={p['ok']} dloss {p['max_abs_dloss']:.2e} dgn {p['max_rel_dgnorm']:.2e} rs_bound {c['rs_err_over_fp32_accumulate_bound']:.3f} ulp {c['rs_max_bf16_ulp_vs_fp32_allreduce']} ag {c['ag_bit_exact']}")
except Exception as e:
    print("FAILED", sys.argv[1:], e); import subprocess; print(subprocess.run(f"grep -n 'File \"/\\|Error' gpurun_out/r2_n8b_{sys.argv[1]}.err | grep -v 'frame #' | tail -12", shell=True, capture_output=True, text=True).stdout[:3000])
PY
done

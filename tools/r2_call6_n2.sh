#!/bin/bash
# two GPUs: scalar all-reduce + loopback tests, dist tests of the symmetric-memory paths, deep-queue bench (no host sync for 20 steps)
mkdir -p gpurun_out
O=gpurun_out
echo skip loopback
timeout 300 python -m pytest tests/test_dist_gpu.py -q -m gpu -s -k "nvls" > $O/r2_n2_dist2.log 2>&1; echo "dist tests rc=$?"; grep -E "passed|failed|rror" $O/r2_n2_dist2.log | tail -5
B200_COMM=nvls B200_BENCH_WATCHDOG_S=120 timeout 140 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29641 \
  bench.py --gpus 2 --steps 20 --warmup 3 --no-cpu-baseline > $O/r2_n2_nvls_s20.json 2> $O/r2_n2_nvls_s20.err; echo "bench nvls steps 20 rc=$?"
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/r2_n2_nvls_s20.json")); p = d["parity"]
    print(f"N=2 nvls: {d['ms_per_step']:.2f} ms {d['value']:.0f} tok/s e2e {d['e2e']['value']:.0f} launches {d['gpu_launches']} parity ok={p['ok']} comm {d["config"]["collectives"]} rs_bound {p["collectives"]["rs_err_over_fp32_accumulate_bound"]:.3f}")
except Exception as e:
    print("FAILED", e); import subprocess; print(subprocess.run(["tail","-5","gpurun_out/r2_n2_nvls_s20.err"],capture_output=True,text=True).stdout)
PY

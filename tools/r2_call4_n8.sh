#!/bin/bash
# Round-2 GPU call 4 (eight GPUs): the reference's own FSDP2/DTensor path and ours under the same unmodified recipe at world 8
# (100-step parity + throughput), then bench.py A/B of the collectives data paths, each with its parity block.
#   gpurun --gpus 8 --timeout 900 -- bash tools/r2_call4_n8.sh
mkdir -p gpurun_out
O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29701 tools/ref_gpu_run.py --strategy fsdp2 --config 8b --steps 100 --out $O/ref_8b_n8.json > $O/ref_8b_n8.log 2>&1; echo "ref 8b n8 rc=$?"; tail -2 $O/ref_8b_n8.log
B200_COMM=nvls timeout 300 $TR --master-port 29702 tools/ref_gpu_run.py --strategy b200_sharded --config 8b --steps 100 --out $O/b200_8b_n8.json > $O/b200_8b_n8.log 2>&1; echo "b200 8b n8 rc=$?"; tail -2 $O/b200_8b_n8.log
python tools/ref_gpu_run.py --compare $O/ref_8b_n8.json $O/b200_8b_n8.json --md $O/r2_parity_n8.md | sed -n 1,9p
port=29710
for cfg in "nvls float32 0" "nvls float32 1" "nccl bfloat16 0" "p2p float32 0"; do
  set -- $cfg
  port=$((port+1))
  B200_COMM=$1 B200_REDUCE_DTYPE=$2 B200_WGRAD_STREAM=$3 B200_BENCH_WATCHDOG_S=150 timeout 170 $TR --master-port $port bench.py --gpus 8 --steps 8 --warmup 3 --no-cpu-baseline > $O/r2_n8_$1_$2_wg$3.json 2> $O/r2_n8_$1_$2_wg$3.err
  echo "bench $cfg rc=$?"
  python - "$1" "$2" "$3" <<'PY'
import json, sys
try:
    d = json.load(open(f"gpurun_out/r2_n8_{sys.argv[1]}_{sys.argv[2]}_wg{sys.argv[3]}.json"))
    p = d.get("parity", {})
    c = p.get("collectives", {})
    print(f"comm={sys.argv[1]} reduce={sys.argv[2]} wgrad_stream={sys.argv[3]}: {d['ms_per_step']:.2f} ms {d['value']:.0f} tok/s  e2e {d['e2e']['value']:.0f} gemm {d['roofline']['achieved']:.0f} TF clocks {d['clocks']['sm_mhz']} | parity ok={p.get('ok')} dloss {p.get('max_abs_dloss'):.2e} dgn {p.get('max_rel_dgnorm'):.2e} rs_ulp {c.get('rs_max_bf16_ulp_vs_fp32_allreduce')} inexact {c.get('rs_frac_not_bit_equal'):.3f} ag {c.get('ag_bit_exact')}")
    for e in c.get("rs_mismatch_examples_rank0", [])[:3]: print("   ", e)
except Exception as e:
    print("FAILED", sys.argv[1:], e); import subprocess; print(subprocess.run(["tail","-8",f"gpurun_out/r2_n8_{sys.argv[1]}_{sys.argv[2]}_wg{sys.argv[3]}.err"],capture_output=True,text=True).stdout)
PY
done
# HSDP 2 x 4 over NCCL + symmetric-memory collectives inside each shard group of 4
timeout 100 $TR --master-port 29731 tools/hsdp_check.py > $O/r2_hsdp_2x4.log 2>&1; echo "hsdp 2x4 rc=$?"; tail -3 $O/r2_hsdp_2x4.log

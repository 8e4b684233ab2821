#!/bin/bash
# Round-2 GPU call 2 (two GPUs): the b200_ctx collectives (NVLS / peer loads on symmetric memory) and the NCCL fp32 / bf16 fallbacks.
#   gpurun --gpus 2 --timeout 900 -- bash tools/r2_call2_n2.sh
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi topo -m 2>/dev/null | head -6
timeout 120 python - <<'PY' 2>&1 | tail -15
# probe: does torch's symmetric memory give us peer pointers and an NVLS multicast mapping on this box?
import os, torch, torch.multiprocessing as mp
def w(rank):
    import torch.distributed as dist, torch.distributed._symmetric_memory as sm
    os.environ["MASTER_ADDR"]="127.0.0.1"; os.environ["MASTER_PORT"]="29511"
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=2, device_id=torch.device("cuda", rank))
    try:
        t = sm.empty(1 << 20, dtype=torch.bfloat16, device=f"cuda:{rank}")
        h = sm.rendezvous(t, group=dist.group.WORLD.group_name)
        print(rank, "backend", sm.get_backend(torch.device("cuda", rank)) if hasattr(sm, "get_backend") else "?", "multicast_support", h.has_multicast_support,
              "mc_ptr", hex(h.multicast_ptr), "bufs", [hex(p) for p in h.buffer_ptrs], "offset", getattr(h, "offset", None), "data_ptr", hex(t.data_ptr()), flush=True)
    except Exception as e:
        print(rank, "symm_mem FAILED:", type(e).__name__, e, flush=True)
    dist.barrier(); dist.destroy_process_group()
if __name__ == "__main__":
    mp.spawn(w, nprocs=2)
PY
timeout 200 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "collective or master" > $O/r2_n2_kernels.log 2>&1; echo "loopback/master tests rc=$?"; tail -4 $O/r2_n2_kernels.log
timeout 500 python -m pytest tests/test_dist_gpu.py -q -m gpu -s > $O/r2_n2_dist.log 2>&1; echo "dist tests rc=$?"; grep -E "passed|failed|error|^\{|Error" $O/r2_n2_dist.log | tail -30
port=29610
for cfg in "nvls float32" "nccl float32" "nccl bfloat16" "p2p float32"; do
  set -- $cfg
  port=$((port+1))
  B200_COMM=$1 B200_REDUCE_DTYPE=$2 B200_BENCH_WATCHDOG_S=150 timeout 170 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 \
    --master-addr 127.0.0.1 --master-port $port bench.py --gpus 2 --steps 6 --warmup 3 --no-cpu-baseline > $O/r2_n2_$1_$2.json 2> $O/r2_n2_$1_$2.err
  echo "bench $1 $2 rc=$?"
  python - "$1" "$2" <<'PY'
import json, sys
try:
    d = json.load(open(f"gpurun_out/r2_n2_{sys.argv[1]}_{sys.argv[2]}.json"))
    p = d.get("parity", {})
    print(f"comm={sys.argv[1]} reduce={sys.argv[2]}: {d['ms_per_step']:.2f} ms {d['value']:.0f} tok/s  e2e {d['e2e']['value']:.0f} gemm {d['roofline']['achieved']:.0f} TF clocks {d['clocks']['sm_mhz']} | parity ok={p.get('ok')} dloss {p.get('max_abs_dloss'):.2e} dgn {p.get('max_rel_dgnorm'):.2e} coll {p.get('collectives')}")
except Exception as e:
    print("FAILED", sys.argv[1:], e); import subprocess; print(subprocess.run(["tail","-5",f"gpurun_out/r2_n2_{sys.argv[1]}_{sys.argv[2]}.err"],capture_output=True,text=True).stdout)
PY
done

#!/bin/bash
# Round-2 GPU call 1 (one GPU): default + experimental suites, reference-vs-ours under the unmodified recipe (tiny, hd128, 8B x 100 steps),
# then the bench A/Bs written at the end of round 1.      gpurun --timeout 1500 -- bash tools/r2_call1.sh
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv,noheader
python -c "import os; print('cpus', len(os.sched_getaffinity(0)))"
timeout 500 python -m pytest tests -q -m gpu -x > $O/r2_gpu_suite.log 2>&1; echo "default gpu suite rc=$?"; tail -3 $O/r2_gpu_suite.log
B200_TEST_EXPERIMENTAL=1 timeout 300 python -m pytest tests/test_experimental_gpu.py -q -m gpu > $O/r2_experimental.log 2>&1; echo "experimental rc=$?"; tail -5 $O/r2_experimental.log

DYN=""
timeout 300 python tools/ref_gpu_run.py --strategy fsdp2 --config tiny --steps 20 --out $O/ref_tiny.json > $O/ref_tiny.log 2>&1
rc=$?; echo "ref tiny rc=$rc"; tail -2 $O/ref_tiny.log
if [ $rc -ne 0 ]; then
  echo "retrying the reference with TORCHDYNAMO_DISABLE=1"; DYN="TORCHDYNAMO_DISABLE=1"
  env $DYN timeout 300 python tools/ref_gpu_run.py --strategy fsdp2 --config tiny --steps 20 --out $O/ref_tiny.json > $O/ref_tiny2.log 2>&1; echo "ref tiny (no dynamo) rc=$?"; tail -2 $O/ref_tiny2.log
fi
timeout 300 python tools/ref_gpu_run.py --strategy b200_sharded --config tiny --steps 20 --out $O/b200_tiny.json > $O/b200_tiny.log 2>&1; echo "b200 tiny rc=$?"; tail -2 $O/b200_tiny.log
timeout 300 python tools/ref_gpu_run.py --strategy b200_sharded --loss fused --config tiny --steps 20 --out $O/b200_tiny_fused.json > $O/b200_tiny_fused.log 2>&1; echo "b200 tiny fused rc=$?"; tail -2 $O/b200_tiny_fused.log
python tools/ref_gpu_run.py --compare $O/ref_tiny.json $O/b200_tiny.json --md $O/r2_parity.md | tail -4
python tools/ref_gpu_run.py --compare $O/ref_tiny.json $O/b200_tiny_fused.json --md $O/r2_parity.md | tail -2

env $DYN timeout 300 python tools/ref_gpu_run.py --strategy fsdp2 --config hd128 --steps 100 --out $O/ref_hd128.json > $O/ref_hd128.log 2>&1; echo "ref hd128 rc=$?"; tail -1 $O/ref_hd128.log
timeout 300 python tools/ref_gpu_run.py --strategy b200_sharded --config hd128 --steps 100 --out $O/b200_hd128.json > $O/b200_hd128.log 2>&1; echo "b200 hd128 rc=$?"; tail -1 $O/b200_hd128.log
python tools/ref_gpu_run.py --compare $O/ref_hd128.json $O/b200_hd128.json --md $O/r2_parity.md | tail -4

env $DYN timeout 400 python tools/ref_gpu_run.py --strategy fsdp2 --config 8b --steps 100 --out $O/ref_8b_n1.json > $O/ref_8b_n1.log 2>&1; echo "ref 8b rc=$?"; tail -2 $O/ref_8b_n1.log
timeout 300 python tools/ref_gpu_run.py --strategy b200_sharded --config 8b --steps 100 --out $O/b200_8b_n1.json > $O/b200_8b_n1.log 2>&1; echo "b200 8b rc=$?"; tail -2 $O/b200_8b_n1.log
python tools/ref_gpu_run.py --compare $O/ref_8b_n1.json $O/b200_8b_n1.json --md $O/r2_parity.md | tail -22

for cfg in "0 engine" "1 engine" "0 facade" "0 engine-swiglu"; do
  set -- $cfg
  fuse=0; api=$2
  if [ "$2" = "engine-swiglu" ]; then fuse=1; api=engine; fi
  B200_FUSE_SWIGLU=$fuse B200_GEMM_SCHED=$1 timeout 200 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --e2e-api $api > $O/r2_n1_sched$1_$2.json 2> $O/r2_n1_sched$1_$2.err
  python - "$1" "$2" <<'PY'
import json, sys
try:
    d = json.load(open(f"gpurun_out/r2_n1_sched{sys.argv[1]}_{sys.argv[2]}.json"))
    print(f"gemm_sched={sys.argv[1]} e2e_api={sys.argv[2]}: {d['ms_per_step']:.2f} ms {d['value']:.0f} tok/s  e2e {d['e2e']['value']:.0f}  gemm {d['roofline']['achieved']:.0f} TF  clocks {d['clocks']['sm_mhz']}")
except Exception as e:
    print("FAILED", sys.argv[1:], e)
PY
done

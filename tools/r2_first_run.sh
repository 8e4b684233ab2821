#!/bin/bash
# First GPU call of the next round (one GPU, ~3 min): everything written after the round-1 GPU budget ran out.
#   gpurun --timeout 600 -- bash tools/r2_first_run.sh
mkdir -p gpurun_out
timeout 400 python -m pytest tests -q -m gpu > gpurun_out/r2_gpu_suite.log 2>&1; echo "default gpu suite rc=$?"; tail -3 gpurun_out/r2_gpu_suite.log
B200_TEST_EXPERIMENTAL=1 timeout 300 python -m pytest tests/test_experimental_gpu.py -q -m gpu > gpurun_out/r2_experimental.log 2>&1; echo "experimental rc=$?"; tail -3 gpurun_out/r2_experimental.log
for cfg in "0 engine" "1 engine" "0 facade" "0 engine-swiglu"; do
  set -- $cfg
  fuse=0; api=$2
  if [ "$2" = "engine-swiglu" ]; then fuse=1; api=engine; fi
  B200_FUSE_SWIGLU=$fuse B200_GEMM_SCHED=$1 timeout 200 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --e2e-api $api > gpurun_out/r2_n1_sched$1_$2.json 2> gpurun_out/r2_n1_sched$1_$2.err
  python - "$1" "$2" <<'PY'
import json, sys
try:
    d = json.load(open(f"gpurun_out/r2_n1_sched{sys.argv[1]}_{sys.argv[2]}.json"))
    print(f"gemm_sched={sys.argv[1]} e2e_api={sys.argv[2]}: {d['ms_per_step']:.2f} ms {d['value']:.0f} tok/s  e2e {d['e2e']['value']:.0f}  gemm {d['roofline']['achieved']:.0f} TF  clocks {d['clocks']['sm_mhz']}")
except Exception as e:
    print("FAILED", sys.argv[1:], e)
PY
done

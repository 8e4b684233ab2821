#!/bin/bash
# Round-end verification on one GPU: full GPU test suite, smoke, the bench line, and the ncu launch list of one profiled step.
mkdir -p gpurun_out
timeout 600 python -m pytest tests -x -q -m gpu > gpurun_out/final_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/final_tests.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/final_smoke.log
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; echo "bench rc=$?"; cut -c1-1500 gpurun_out/final_bench.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 1200 --csv --log-file gpurun_out/final_launches.csv python bench.py --profile --steps 1 > gpurun_out/final_prof.log 2>&1; echo "ncu rc=$?"

#!/bin/bash
# Debug: rebuild libb200_train.so with in-kernel phase timers for the attention backward and print cycles per (q tile, kv tile) pair.
set -e
cd "$(dirname "$0")/../automodel_b200/csrc"
for f in gemm_tcgen05 elementwise attention attention_tc c_api; do
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -DB200_ATTN_PROFILE -c $f.cu -o /tmp/prof_$f.o &
done
wait
nvcc -shared -o /tmp/libb200_prof.so /tmp/prof_*.o -L/usr/local/cuda/lib64 -lcublasLt
cp libb200_train.so /tmp/libb200_train.bak; cp /tmp/libb200_prof.so libb200_train.so
cd ../..
python - <<'PY'
import torch, sys
sys.path.insert(0, ".")
from automodel_b200 import ops
T, Hq, Hkv, D = 4096, 32, 8, 128
qkv = torch.randn(T, (Hq + 2 * Hkv) * D, device="cuda").bfloat16()
cu = torch.tensor([0, T], dtype=torch.int32, device="cuda")
q, k, v = qkv[:, :Hq * D], qkv[:, Hq * D:(Hq + Hkv) * D], qkv[:, (Hq + Hkv) * D:]
o, lse = ops.attn_fwd(q, k, v, cu, T, Hq, Hkv, D)
do = torch.randn_like(o); dqkv = torch.empty_like(qkv)
for i in range(2):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    ops.attn_bwd(q, k, v, o, do, lse, cu, T, Hq, Hkv, D, dqkv[:, :Hq * D], dqkv[:, Hq * D:(Hq + Hkv) * D], dqkv[:, (Hq + Hkv) * D:])
    e1.record(); torch.cuda.synchronize()
    print("attn_bwd total ms", e0.elapsed_time(e1), file=sys.stderr)
PY
cp /tmp/libb200_train.bak automodel_b200/csrc/libb200_train.so

#!/usr/bin/env python
"""HSDP on real GPUs (reference: distributed.dp_replicate_size, mesh (dp_replicate, dp_shard), components/distributed/mesh_utils.py:116-190):
8 ranks as 2 replicas x 4 shards - reduce-scatter inside each shard group, gradient-shard all-reduce across the replicas - against one rank
accumulating all 8 sequences.  torchrun --nproc-per-node 8 tools/hsdp_check.py"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from automodel_b200 import diagnostics as D

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
R = 2
S = world // R
shard_groups = [dist.new_group(list(range(r * S, (r + 1) * S))) for r in range(R)]
replica_groups = [dist.new_group([s + r * S for r in range(R)]) for s in range(S)]
pg, rpg = shard_groups[rank // S], replica_groups[rank % S]
out = {}
for comm in ("nccl", "nvls"):
    res = D.check_sharded_step_parity(pg, dev, steps=10, replica_group=rpg, comm=comm, reduce_dtype="float32")
    out[comm] = res
    ok = res["ranks_agree"] and res["max_abs_dloss"] <= 1e-3 and res["max_rel_dgnorm"] <= 2e-2
    if rank == 0:
        print(f"HSDP {R}x{S} comm={comm}: ok={ok} {json.dumps(res)}", flush=True)
dist.barrier()
dist.destroy_process_group()

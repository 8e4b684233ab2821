"""Multi-GPU parity of the sharded step (NCCL, world size 2): in-place reduce-scatter / all-gather on the flat buffers.
Skipped on a single-GPU box; the same orchestration is covered on CPU over gloo in tests/test_engine_cpu.py."""
import os
import pytest
import torch
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu

from tests.golden_utils import load, model_cfg, init_params, batches  # noqa: E402


def _worker(rank, world, port, out_q, peer_comm="1"):
    os.environ["B200_PEER_COMM"] = peer_comm
    import torch.distributed as dist
    from automodel_b200.engine import ShardedLlamaEngine
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    z, meta = load("tiny_bf16")
    cfg = model_cfg(meta); oc = meta["optimizer"]
    eng = ShardedLlamaEngine(cfg, f"cuda:{rank}", process_group=dist.group.WORLD, max_tokens=meta["config"]["seq"], lr=oc["lr"],
                             betas=tuple(oc["betas"]), eps=oc["eps"], weight_decay=oc["weight_decay"], adam_mode=1)
    eng.load_state_dict(init_params(meta))
    res = []
    for s in range(10):
        b = batches(z, meta, s)[0]
        mb = {"input_ids": torch.from_numpy(b["input_ids"][rank:rank + 1]), "labels": torch.from_numpy(b["labels"][rank:rank + 1])}
        l, g = eng.train_step([mb], meta["max_grad_norm"])
        res.append((float(l), float(g)))
    flat = torch.cat([p.float().reshape(-1) for p in eng.state_dict().values()])
    gathered = [torch.empty_like(flat) for _ in range(world)]
    dist.all_gather(gathered, flat)
    same = all(torch.equal(gathered[0], t) for t in gathered)
    if rank == 0:
        out_q.put((res, same))
    dist.barrier()
    dist.destroy_process_group()


_MODES = [("1", "nvlink_peer_path"), ("0", "nccl_collectives")]
if os.environ.get("B200_TEST_EXPERIMENTAL") == "1":     # written after the last multi-GPU run of round 1: opt-in until it has passed once
    _MODES.append(("ag", "nccl_rs_copy_engine_ag"))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("peer_comm", [m for m, _ in _MODES], ids=[i for _, i in _MODES])
def test_world2_matches_reference_curve(peer_comm):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29700 + (os.getpid() % 1000) + (7 if peer_comm == "0" else 0)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q, peer_comm)) for r in range(2)]
    for p in procs:
        p.start()
    res, same = q.get(timeout=600)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert same, "ranks disagree on the gathered parameters after the in-place all-gather"
    _, meta = load("tiny_bf16")
    for s, (l, g) in enumerate(res):
        assert abs(l - meta["loss"][s]) < 1e-3, (s, l, meta["loss"][s])
        assert abs(g - meta["grad_norm"][s]) < 2e-2 * meta["grad_norm"][s], (s, g, meta["grad_norm"][s])

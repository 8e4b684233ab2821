import sys, time, cProfile, pstats, torch
sys.path.insert(0, ".")
import bench
from automodel_b200.engine import ShardedLlamaEngine
cfg = dict(bench.LLAMA3_8B); cfg["num_hidden_layers"] = 8
eng = ShardedLlamaEngine(cfg, "cuda", max_tokens=4096, adam_mode=1, max_positions=4096)
eng.init_random_(0)
ids = torch.randint(0, 128256, (1, 4096)); lab = torch.full_like(ids, -100); lab[:, :-1] = ids[:, 1:]
st = [eng.stage(ids, lab)]
for _ in range(3): eng.train_step(None, 1.0, num_label_tokens=4095, staged=st)
torch.cuda.synchronize()
t = time.perf_counter(); eng.train_step(None, 1.0, num_label_tokens=4095, staged=st); h = time.perf_counter() - t
torch.cuda.synchronize(); print("host enqueue ms (8 layers):", h * 1e3)
pr = cProfile.Profile(); pr.enable()
eng.train_step(None, 1.0, num_label_tokens=4095, staged=st)
pr.disable(); torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("tottime").print_stats(14)

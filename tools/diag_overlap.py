"""Diagnostic: is the first-step loss a function of the stream-overlap switch or of engine construction order?"""
import sys, torch
sys.path.insert(0, ".")
from automodel_b200.engine import ShardedLlamaEngine
cfg = {"vocab_size": 32768, "hidden_size": 4096, "intermediate_size": 14336, "num_hidden_layers": 4, "num_attention_heads": 32,
       "num_key_value_heads": 8, "max_position_embeddings": 8192, "rms_norm_eps": 1e-5, "rope_theta": 500000.0}
g = torch.Generator().manual_seed(3)
ids = torch.randint(0, 32768, (1, 4096), generator=g)
lab = torch.full_like(ids, -100); lab[:, :-1] = ids[:, 1:]
for on in (False, True, False, True):
    eng = ShardedLlamaEngine(cfg, "cuda", max_tokens=4096, lr=1e-5, adam_mode=1, max_positions=4096)
    eng.init_random_(seed=5)
    csum = float(sum(p.float().sum() for p in eng.p_full))
    eng.set_stream_overlap(on)
    eng.loss_dev.zero_(); eng.forward_backward(ids, lab, None, 4095); torch.cuda.synchronize()
    l_a = float(eng.loss_dev[0]); gsum_a = float(sum(x.float().abs().sum() for x in eng.g_full))
    eng.loss_dev.zero_(); eng.forward_backward(ids, lab, None, 4095); torch.cuda.synchronize()
    l_b = float(eng.loss_dev[0]); gsum_b = float(sum(x.float().abs().sum() for x in eng.g_full))
    loss, gn = eng.train_step([{"input_ids": ids, "labels": lab}], 1.0)
    print(f"overlap={on}: param checksum {csum!r}  fwd/bwd loss {l_a!r} {l_b!r}  |g| sum {gsum_a!r} {gsum_b!r}  train_step loss {float(loss)!r} gn {float(gn)!r}", flush=True)
    del eng; torch.cuda.empty_cache()

"""Full-size check of the decode graph with the forked L2 weight prefetch: same tokens as without it, and timing per
prefetch budget.  Usage: python scripts/decode_prefetch_check.py [MB ...]"""
import os
import sys
import time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from videollama2_b200 import presets
from videollama2_b200.model import VLLMs

dev = torch.device("cuda:0")
cfg = presets.make_config(presets.MISTRAL_7B, 16)
model = VLLMs[cfg.model_type].from_state_dict(cfg, presets.random_state_dict(cfg, dev), device=dev)
px = torch.randn((16, 3, 336, 336), generator=torch.Generator().manual_seed(1)).to(torch.bfloat16).to(dev)
ids = torch.randint(3, cfg.vocab_size, (1, 256), generator=torch.Generator().manual_seed(2))
ids[0, 4] = -201
dec = model.get_model().decoder


def run(n):
    torch.cuda.synchronize()
    t = time.perf_counter()
    out = model.generate(ids, images=[(px, "video")], max_new_tokens=n, do_sample=False, use_cache=True, eos_token_id=None)
    torch.cuda.synchronize()
    return out, (time.perf_counter() - t) * 1e3


N = int(os.environ.get('N_TOK', '40'))
ref, _ = run(N)                      # eager decode loop
budgets = [float(a) for a in sys.argv[1:]] or [0, 34]
for mb in budgets:
    dec.decode_prefetch_mb = mb
    model.enable_cuda_graphs(True)
    run(4)
    a, ta = run(3)
    b, tb = run(N)
    same = bool(torch.equal(b, ref))
    # logits of the last replay against the eager loop's
    print(f"prefetch {mb:6.1f} MB: {(tb - ta) / (N - 3):.3f} ms/token, tokens equal eager: {same}, first 8 {b[0, :8].tolist()}")
    model.enable_cuda_graphs(False)

"""Per-op CUDA-event timing of the ViT tower for a given number of frames (how the frame-sharded shard behaves)."""
import collections
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from videollama2_b200 import ops, presets
from videollama2_b200.model import encoder as enc_mod
from videollama2_b200.model.encoder import CLIPVisionTower

frames = int(sys.argv[1]) if len(sys.argv) > 1 else 2
dev = torch.device("cuda")
cfg = presets.make_config(presets.MISTRAL_7B, 16)
sd = {k: v for k, v in presets.random_state_dict(cfg, dev).items() if "vision_tower" in k}
tower = CLIPVisionTower("synthetic-clip", cfg, vision_config=cfg.vision_config).load_state_dict(
    sd, dev, prefix="model.vision_tower.vision_tower.vision_model.")
px = torch.randn((frames, 3, 336, 336), device=dev).bfloat16()
for _ in range(3):
    tower(px)
recs = collections.OrderedDict()


class Proxy:
    def __getattr__(self, name):
        fn = getattr(ops, name)
        if not callable(fn) or name.startswith("ACT"):
            return fn

        def wrapped(*a, **k):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            out = fn(*a, **k)
            e1.record()
            shp = tuple(a[0].shape) if a and hasattr(a[0], "shape") else ()
            w = tuple(a[1].shape) if len(a) > 1 and hasattr(a[1], "shape") else ()
            recs.setdefault((name, shp, w), []).append((e0, e1))
            return out
        return wrapped


enc_mod.ops = Proxy()
tower(px)
torch.cuda.synchronize()
enc_mod.ops = ops
tot = 0.0
for (name, shp, w), evs in recs.items():
    ms = sum(a.elapsed_time(b) for a, b in evs)
    tot += ms
    print(f"{name:18s} {str(shp):22s} {str(w):16s} n={len(evs):3d} total={ms * 1e3:8.1f}us avg={ms * 1e3 / len(evs):7.1f}us")
print(f"frames={frames} sum={tot:.3f} ms")
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
tower.enable_cuda_graphs(True)
for _ in range(3):
    tower(px)
e0.record()
for _ in range(10):
    tower(px)
e1.record()
torch.cuda.synchronize()
print(f"graph replay: {e0.elapsed_time(e1) / 10:.3f} ms per call")

"""One rank of the multi-GPU frame-parallel check (launched by torchrun, NCCL): the frame-sharded ViT + all-gather must
reproduce the single-GPU tower output BIT-EXACTLY (frame sharding must not change numerics)."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    rank, world = dist.get_rank(), dist.get_world_size()
    from helpers import build_engine
    from oracle import synth
    from videollama2_b200 import parallel
    ok = True
    for name, frames in (("mid", 6), ("mid", 7), ("tiny", 1)):
        cfg = synth.CONFIGS[name]
        sd = synth.state_dict(cfg)
        model = build_engine(cfg, sd, dev)
        g = torch.Generator().manual_seed(99)
        px = torch.randn((frames, 3, cfg.vision.image, cfg.vision.image), generator=g).to(torch.bfloat16).to(dev)
        tower = model.get_vision_tower()
        full = tower(px)
        got = parallel.encode_frames_sharded(tower, px)
        ok &= bool(torch.equal(got, full))
        # epilogue-fused gather (peer / multicast stores from the last ViT GEMM) gives the same bits
        for mcast in (False, True):
            fg = parallel.FusedFrameGather(tower, frames, use_multicast=mcast)
            for _ in range(2):                      # twice: buffer reuse across calls
                ok &= bool(torch.equal(fg.encode(px), full))
        # and the connector consumes the gathered tokens identically
        a = model.get_model().mm_projector(got[None])
        b = model.get_model().mm_projector(full[None])
        ok &= bool(torch.equal(a, b))
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    dist.destroy_process_group()
    print(f"RANK{rank} {'OK' if int(flag.item()) == 1 else 'FAIL'}", flush=True)
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()

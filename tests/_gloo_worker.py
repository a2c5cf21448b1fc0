"""Worker for tests/test_parallel_gloo.py: one rank of a world_size-N gloo job on CPU (RANK/WORLD_SIZE/MASTER_* in env)."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


class FakeTower:
    """CPU stand-in with the tower's call surface: per-frame deterministic features (bit-exact under any sharding)."""
    num_patches, hidden_size = 5, 8

    def __call__(self, frames):
        base = frames.float().mean(dim=(1, 2, 3))
        return (base[:, None, None] + torch.arange(5.)[None, :, None] * 0.5 + torch.arange(8.)[None, None, :]).to(frames.dtype)


class FakeProjector:
    """CPU stand-in for the STC connector's two halves: a per-frame map (s1) and a part that mixes neighbouring frames."""
    hidden_size = 8
    s1_dtype = torch.float32

    def forward_s1(self, x):                       # [f,h,w,C] -> [f,h,w,C], per frame (mean over the frame's pixels)
        return x * 2 + x.mean(dim=(1, 2), keepdim=True)

    def forward_from_s1(self, a):                  # [b,T,h,w,C] -> [b, T*h*w, C]; couples frame t with t-1
        prev = torch.cat([torch.zeros_like(a[:, :1]), a[:, :-1]], 1)
        return (a + 0.5 * prev).flatten(1, 3)

    def __call__(self, feats):                     # the unsharded connector: [b,T,n,C]
        b, T, n, C = feats.shape
        hw = int(n ** 0.5)
        return self.forward_from_s1(self.forward_s1(feats.reshape(b * T, hw, hw, C)).view(b, T, hw, hw, C))


class FakeModel:
    def __init__(self, tower, proj):
        self.vision_tower, self.mm_projector = tower, proj

    def get_model(self):
        return self

    def get_vision_tower(self):
        return self.vision_tower

    def temporal_aggregator(self, feats):
        return self.mm_projector(feats)


class SquareTower(FakeTower):
    num_patches = 4                                 # 2 x 2 grid so that the connector sees a square frame

    def __call__(self, frames):
        return super().__call__(frames)[:, :4]


def check_frame_parallel(F):
    """parallel.FrameParallel.encode (the product path behind encode_images_or_videos) == the unsharded computation,
    with and without the first connector stage sharded, for one video and for a batch of two."""
    from videollama2_b200.parallel import FrameParallel
    ok = True
    for b in (1, 2):
        if F % b:
            continue
        t = F // b
        frames = torch.randn((F, 3, 4, 4), generator=torch.Generator().manual_seed(3))
        model = FakeModel(SquareTower(), FakeProjector())
        want = model.mm_projector(model.vision_tower(frames).view(b, t, 4, 8))
        for shard_s1 in (True, False):
            got = FrameParallel(None, shard_s1=shard_s1).encode(model, frames, b, t)
            ok = ok and torch.equal(got, want)
    return ok


def main():
    F = int(sys.argv[1])
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    from videollama2_b200.parallel import all_gather_frames, encode_frames_sharded, frame_shard
    frames = torch.randn((F, 3, 4, 4), generator=torch.Generator().manual_seed(0))
    tower = FakeTower()
    full = tower(frames)
    got = encode_frames_sharded(tower, frames)
    a, b = frame_shard(F, rank, world)
    ok = torch.equal(got, full) and torch.equal(all_gather_frames(full[a:b].clone(), F), full)
    try:
        all_gather_frames(torch.zeros((b - a + 3, 5, 8)), F)   # wrong shard size must be refused before any collective
        ok = False
    except ValueError:
        pass
    ok = ok and check_frame_parallel(F)
    dist.barrier()
    dist.destroy_process_group()
    print(f"RANK{rank} {'OK' if ok else 'FAIL'}", flush=True)
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()

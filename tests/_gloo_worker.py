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
    dist.barrier()
    dist.destroy_process_group()
    print(f"RANK{rank} {'OK' if ok else 'FAIL'}", flush=True)
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()

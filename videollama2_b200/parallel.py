"""Frame-parallel vision stage: the ViT treats frames as batch (videollama2_arch.py:130-131), so one video's frames are
sharded over the ranks of a node, encoded independently, and exchanged with ONE all-gather of visual tokens
(NCCL over NVLink 5 / NVSwitch) before the first op that mixes time (the connector's Conv3d, projector.py:208).
The reference has no such path (its multi-GPU inference is one process per GPU over dataset chunks,
scripts/eval/eval_video_mcqa_mvbench.sh:8-34); this is the exchange step BASELINE.json's north_star names.

Host logic is backend-agnostic (tests run it with gloo, world_size 2, on CPU tensors)."""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch
import torch.distributed as dist


def frame_shard(num_frames: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [start, end) of frames for `rank`; sizes differ by at most one (first ranks take the remainder)."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world {world}")
    base, rem = divmod(num_frames, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def shard_sizes(num_frames: int, world: int) -> List[int]:
    return [frame_shard(num_frames, r, world)[1] - frame_shard(num_frames, r, world)[0] for r in range(world)]


def all_gather_frames(local: torch.Tensor, num_frames: int, group=None) -> torch.Tensor:
    """local [f_r, n, C] on each rank -> [num_frames, n, C] on every rank, frame order preserved."""
    world = dist.get_world_size(group)
    sizes = shard_sizes(num_frames, world)
    rank = dist.get_rank(group)
    if local.shape[0] != sizes[rank]:
        raise ValueError(f"rank {rank} holds {local.shape[0]} frames, expected {sizes[rank]}")
    out = torch.empty((num_frames,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    if len(set(sizes)) == 1 and sizes[0] > 0:
        dist.all_gather_into_tensor(out, local.contiguous(), group=group)
        return out
    # ragged: pad every shard to the largest one, gather, then compact
    mx = max(sizes)
    pad = torch.zeros((mx,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    buf = torch.empty((world * mx,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(buf, pad, group=group)
    pos = 0
    for r, n in enumerate(sizes):
        out[pos:pos + n] = buf[r * mx: r * mx + n]
        pos += n
    return out


def encode_frames_sharded(tower, frames: torch.Tensor, group=None) -> torch.Tensor:
    """frames [F,3,H,W] (identical on every rank) -> [F, n, C] on every rank: local ViT on this rank's shard + all-gather."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    a, b = frame_shard(frames.shape[0], rank, world)
    if b > a:
        local = tower(frames[a:b])
    else:
        n = tower.num_patches
        local = torch.empty((0, n, tower.hidden_size), dtype=frames.dtype, device=frames.device)
    return all_gather_frames(local, frames.shape[0], group)


class FrameParallel:
    """The frame-parallel vision stage as a PRODUCT path (north_star: "per-frame ViT encode shards across the GPUs of one
    box with an all-gather of visual tokens before the connector, the LLM running on rank 0").

    Every rank of `group` calls `model.encode_images_or_videos(images)` / `generate(...)` with IDENTICAL inputs (SPMD).
    Rank r encodes frames [a_r, b_r) of the flattened (video, frame) batch with the ViT and - `shard_s1` - the
    connector's first RegStage, which is per-frame too (projector.py:203-205; SE pools per frame), i.e. everything in front
    of the first op that mixes time (the Conv3d, projector.py:208).  ONE all-gather ([F, 576, 4096] bf16 with s1 sharded,
    [F, 576, 1024] without) then puts the full tensor on every rank; sampler + s2 + readout run on what was gathered.
    Frame sharding never changes a GEMM's K-order, so the result is bit-identical to the single-GPU path
    (tests/test_multigpu_gpu.py).  `llm_rank`: the rank whose generate() / forward() goes on to the decoder; the other
    ranks return None right after the collective."""

    def __init__(self, group=None, shard_s1: bool = True, llm_rank: int = 0):
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.shard_s1 = shard_s1
        self.llm_rank = llm_rank

    def encode(self, model, frames: torch.Tensor, b: int, t: int) -> torch.Tensor:
        """frames [(b t),3,H,W] identical on every rank -> connector output [b, L, D] on every rank."""
        tower = model.get_model().get_vision_tower()
        proj = model.get_model().mm_projector
        F = frames.shape[0]
        a, e = frame_shard(F, self.rank, self.world)
        n, C = tower.num_patches, tower.hidden_size
        local = tower(frames[a:e]) if e > a else torch.empty((0, n, C), dtype=frames.dtype, device=frames.device)
        if self.shard_s1 and hasattr(proj, "forward_from_s1"):
            hw = int(n ** 0.5)
            H = proj.hidden_size
            if e > a:
                s1 = proj.forward_s1(local.view(e - a, hw, hw, C)).reshape(e - a, n, H)
            else:
                s1 = torch.empty((0, n, H), dtype=proj.s1_dtype, device=frames.device)   # more ranks than frames
            full = all_gather_frames(s1, F, self.group)
            return proj.forward_from_s1(full.view(b, t, hw, hw, H)).to(frames.dtype)
        feats = all_gather_frames(local, F, self.group)
        return model.temporal_aggregator(feats.view(b, t, n, C))


class NvlsAllReduce:
    """All-reduce of the tensor-parallel decoder's partial sums as ONE kernel over NVLink / NVSwitch (csrc/tp_allreduce.cu):
    reduction of the ranks' bf16 partials with peer loads (fp32 accumulation in rank order: bit-identical to NCCL's result;
    `inswitch_reduce=True` uses multimem.ld_reduce instead, whose bf16 rounding differs - see the kernel), the row sums of
    squares the next folded RMSNorm needs, and the broadcast of both through the switch (multimem.st) - with the cross-GPU
    barriers inside the kernel.
    Buffers live in symmetric memory (torch.distributed._symmetric_memory): `part` is where the row-parallel GEMM writes
    its output, `out[i]` (two, alternating) receive the reduced stream.  Falls back to peer loads / stores when the
    allocation has no multicast mapping."""

    def __init__(self, max_rows: int, hidden: int, device, group=None, use_multicast: bool = True,
                 inswitch_reduce: bool = False, dtype: torch.dtype = torch.bfloat16):
        import torch.distributed._symmetric_memory as symm_mem
        self.inswitch_reduce = inswitch_reduce
        self.group = group if group is not None else dist.group.WORLD
        self.rank = dist.get_rank(self.group)
        self.world = dist.get_world_size(self.group)
        if self.world > 8:
            raise ValueError("NvlsAllReduce supports up to 8 ranks (one NVSwitch domain)")
        self.max_rows, self.hidden = max_rows, hidden
        dev = torch.device(device)

        def sym(shape, dtype):
            t = symm_mem.empty(shape, dtype=dtype, device=dev)
            t.zero_()
            return t, symm_mem.rendezvous(t, self.group)

        self.dtype = dtype
        self.part, self._h_part = sym((max_rows, hidden), dtype)
        self.out = []
        self._h_out = []
        self.stats = []
        self._h_stats = []
        for _ in range(2):
            t, h = sym((max_rows, hidden), dtype)
            self.out.append(t)
            self._h_out.append(h)
            t, h = sym((max_rows,), torch.float32)
            self.stats.append(t)
            self._h_stats.append(h)
        self.pads, self._h_pads = sym((32,), torch.int32)
        self.multicast = False
        if use_multicast:
            try:
                self.multicast = all(int(h.multicast_ptr) != 0 for h in [self._h_part] + self._h_out + self._h_stats)
            except Exception:
                self.multicast = False
        torch.cuda.synchronize(dev)
        dist.barrier(self.group)        # every rank's pads are zeroed before anybody signals
        self.epoch = 0
        self.turn = 0

    def reduce(self, rows: int):
        """Sum `self.part[:rows]` over the ranks.  Returns (x [rows, H] bf16, sumsq [rows, 1] fp32), valid until the call
        after next (the two output buffers alternate)."""
        from . import _lib
        import ctypes as C
        if rows > self.max_rows:
            raise ValueError(f"{rows} rows exceed the symmetric buffers ({self.max_rows})")
        i = self.turn
        self.turn ^= 1
        self.epoch += 1
        a = _lib.TpAllReduceArgs(rank=self.rank, world=self.world, S=rows, H=self.hidden, epoch=self.epoch)
        for r in range(self.world):
            a.part[r] = int(self._h_part.buffer_ptrs[r])
            a.xout[r] = int(self._h_out[i].buffer_ptrs[r])
            a.stats[r] = int(self._h_stats[i].buffer_ptrs[r])
            a.pads[r] = int(self._h_pads.buffer_ptrs[r])
        if self.multicast:
            a.part_mc = int(self._h_part.multicast_ptr)
            a.xout_mc = int(self._h_out[i].multicast_ptr)
            a.stats_mc = int(self._h_stats[i].multicast_ptr)
            a.inswitch_reduce = 1 if self.inswitch_reduce else 0
        _lib.check(_lib.load(self.dtype).vl2_tp_allreduce_stats(C.byref(a), torch.cuda.current_stream().cuda_stream),
                   "vl2_tp_allreduce_stats")
        return self.out[i][:rows], self.stats[i][:rows].view(rows, 1)

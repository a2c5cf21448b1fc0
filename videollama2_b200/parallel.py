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


class FusedFrameGather:
    """Frame-parallel ViT whose LAST GEMM writes its output tiles straight into every rank's gather buffer.

    The buffer [F*(np+1), C] lives in symmetric memory (torch.distributed._symmetric_memory: same allocation mapped
    into every process of the node over NVLink/NVSwitch).  Rank r computes frames [a_r, b_r) and its final fc2 GEMM
    (vl2_gemm_bf16 with bcast_out / mc_out) stores each output vector to its own rows of the local buffer AND to the
    same rows of the peers' buffers (P2P stores, or one multimem.st through the switch when multicast is available),
    tile by tile while the remaining tiles are still being computed.  A symmetric-memory barrier then stands in for the
    collective's completion.  No NCCL kernel, no separate copy pass."""

    def __init__(self, tower, num_frames: int, group=None, use_multicast: bool = True):
        import torch.distributed._symmetric_memory as symm_mem
        self.tower = tower
        self.group = group if group is not None else dist.group.WORLD
        self.rank = dist.get_rank(self.group)
        self.world = dist.get_world_size(self.group)
        self.F = num_frames
        self.S = tower.seq_len              # tokens per frame in the residual stream (CLIP: patches + CLS)
        self.C = tower.hidden_size
        dev = tower.device
        self.buf = symm_mem.empty((num_frames * self.S, self.C), dtype=torch.bfloat16, device=dev)
        self.hdl = symm_mem.rendezvous(self.buf, self.group)
        self.ptrs = [int(p) for p in self.hdl.buffer_ptrs]
        mc = 0
        if use_multicast:
            try:
                mc = int(self.hdl.multicast_ptr) if self.hdl.has_multicast_support(dev.type, dev.index) else 0
            except Exception:
                mc = 0
        self.mc_ptr = mc

    def encode(self, frames: torch.Tensor) -> torch.Tensor:
        """frames [F,3,H,W] (identical on every rank) -> [F, np, C] on every rank."""
        if frames.shape[0] != self.F:
            raise ValueError(f"FusedFrameGather was built for {self.F} frames, got {frames.shape[0]}")
        a, b = frame_shard(self.F, self.rank, self.world)
        row_bytes = self.C * 2
        self.hdl.barrier(channel=0)              # every rank finished reading the previous result (buffer reuse)
        if b > a:
            local = self.buf[a * self.S: b * self.S]
            off = a * self.S * row_bytes
            if self.mc_ptr:
                self.tower.hidden_states(frames[a:b].to(torch.bfloat16).contiguous(), last_out=local,
                                         mc_ptr=self.mc_ptr + off)
            else:
                peers = [ptr + off for r, ptr in enumerate(self.ptrs) if r != self.rank]
                self.tower.hidden_states(frames[a:b].to(torch.bfloat16).contiguous(), last_out=local, bcast_ptrs=peers)
        self.hdl.barrier(channel=1)              # all ranks' tiles have landed everywhere
        return self.tower.feature_select(self.buf.view(self.F, self.S, self.C)).contiguous()


def bench_frame_parallel(model, px_dev: torch.Tensor, rank: int, world: int, dev, iters: int = 5,
                         fused: bool = True) -> dict:
    """Device-timed (CUDA events, max over ranks) ViT(shard) + all-gather for one 16-frame video."""
    tower = model.get_vision_tower()
    F = px_dev.shape[0]
    for _ in range(2):
        encode_frames_sharded(tower, px_dev)
    times = []
    for _ in range(iters):
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        a, b = frame_shard(F, rank, world)
        e0.record()
        local = tower(px_dev[a:b])
        e1.record()
        all_gather_frames(local, F)
        e2.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e2), e0.elapsed_time(e1), e1.elapsed_time(e2)], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        times.append(t.tolist())
    times.sort(key=lambda x: x[0])
    tot, vit, gat = times[len(times) // 2]
    if not fused:
        return {"ranks": world, "frames_per_rank": shard_sizes(F, world), "vit_shard_plus_gather_ms": tot,
                "vit_shard_ms": vit, "all_gather_ms": gat, "frames_per_s": F / (tot * 1e-3),
                "gather_bytes": int(F * tower.num_patches * tower.hidden_size * 2)}
    fused = None
    try:
        import os
        fg = FusedFrameGather(tower, F, use_multicast=os.environ.get("VL2_BENCH_MULTICAST", "0") == "1")
        ref = encode_frames_sharded(tower, px_dev)
        got = fg.encode(px_dev)
        exact = bool(torch.equal(got, ref))
        ft = []
        for _ in range(iters):
            dist.barrier()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fg.encode(px_dev)
            e1.record()
            torch.cuda.synchronize()
            t = torch.tensor([e0.elapsed_time(e1)], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ft.append(float(t.item()))
        ft.sort()
        fused = {"ms": ft[len(ft) // 2], "frames_per_s": F / (ft[len(ft) // 2] * 1e-3), "bit_exact_vs_nccl": exact,
                 "multicast": bool(fg.mc_ptr)}
    except Exception as e:  # symmetric memory may be unavailable on a given box: report, do not fail the bench
        fused = {"error": repr(e)[:300]}
    return {"ranks": world, "fused_epilogue_gather": fused, "frames_per_rank": shard_sizes(F, world), "vit_shard_plus_gather_ms": tot, "vit_shard_ms": vit,
            "all_gather_ms": gat, "frames_per_s": F / (tot * 1e-3),
            "gather_bytes": int(F * tower.num_patches * tower.hidden_size * 2)}

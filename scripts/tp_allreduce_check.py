"""2+ GPU check of vl2_tp_allreduce_stats against NCCL all_reduce + vl2_row_sumsq (torchrun --nproc-per-node N)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
rank, world = dist.get_rank(), dist.get_world_size()
from videollama2_b200 import ops
from videollama2_b200.parallel import NvlsAllReduce

for S, H in ((1776, 8192), (139, 256), (38, 512)):
    for mc, insw in ((True, False), (True, True), (False, False)):
        nv = NvlsAllReduce(S, H, dev, None, use_multicast=mc, inswitch_reduce=insw)
        for it in range(3):
            g = torch.Generator(device=dev).manual_seed(100 * rank + it)
            part = torch.randn((S, H), generator=g, device=dev).to(torch.bfloat16)
            ref = part.clone()
            dist.all_reduce(ref)
            ss_ref = ops.row_sumsq(ref)
            nv.part[:S].copy_(part)
            x, ss = nv.reduce(S)
            torch.cuda.synchronize()
            dx = (x.float() - ref.float()).abs().max().item()
            nbad = int((x != ref).sum())
            dss = ((ss - ss_ref).abs() / ss_ref.abs().clamp_min(1e-9)).max().item()
            # exact fp32 reference of the sum
            parts = [torch.empty_like(part) for _ in range(world)]
            dist.all_gather(parts, part)
            exact = sum(p.float() for p in parts).to(torch.bfloat16)
            print(f"rank{rank} S={S} H={H} mc={mc and nv.multicast} inswitch={insw} it={it}: max|x-nccl|={dx:.4g} mismatches={nbad} "
                  f"vs_fp32_sum: ours={int((x != exact).sum())} nccl={int((ref != exact).sum())} stats_rel={dss:.3g}", flush=True)
        del nv
dist.barrier()
dist.destroy_process_group()

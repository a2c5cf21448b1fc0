"""Worker for tests/test_parallel_gloo.py::test_tp_partition_gloo: one rank of a world_size-N gloo job on CPU.

Checks the tensor-parallel partitioning of the decoder (videollama2_b200/model/tp_decoder.py: shard_plan / shard_state_dict)
by running the sharded mathematics in plain fp32 torch on this rank's slices - local attention heads, partial o_proj / down_proj
sums combined with a real all-reduce, vocab-parallel logits combined with a real all-gather - against the unsharded oracle
forward (oracle/torch_ref.py).  The CUDA kernels are not involved: this pins WHICH rows / columns every rank owns."""
import os
import sys

import torch
import torch.distributed as dist
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    name = sys.argv[1]
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    from helpers import engine_config
    from oracle import synth, torch_ref
    from videollama2_b200.model.tp_decoder import shard_plan, shard_state_dict
    cfg = synth.CONFIGS[name]
    l = cfg.llm
    sd = {k: v.float() for k, v in synth.iter_state(synth.llm_specs(l))}
    ec = engine_config(cfg)
    plan = shard_plan(ec, rank, world)
    loc = shard_state_dict(sd, ec, rank, world)
    S = 23
    x = torch.randn((S, l.hidden), generator=torch.Generator().manual_seed(5)) * 0.5
    ref = torch_ref.decoder_forward(sd, l, x, torch.float32, all_logits=True)

    D, hq, hkv = plan["D"], plan["Hq"], plan["Hkv"]
    cos, sin = torch_ref.rope_cos_sin(S, D, l.theta, torch.float32)
    h = x.clone()
    for i in range(l.layers):
        p = f"model.layers.{i}."
        y = torch_ref.rmsnorm(h, loc[p + "input_layernorm.weight"], l.eps)

        def proj(nm):
            b = loc.get(p + f"self_attn.{nm}.bias")
            return F.linear(y, loc[p + f"self_attn.{nm}.weight"], b)

        q = proj("q_proj").view(S, hq, D).transpose(0, 1)
        k = proj("k_proj").view(S, hkv, D).transpose(0, 1)
        v = proj("v_proj").view(S, hkv, D).transpose(0, 1)
        q = q * cos + torch_ref._rot_half(q) * sin
        k = k * cos + torch_ref._rot_half(k) * sin
        k = k.repeat_interleave(hq // hkv, 0)
        v = v.repeat_interleave(hq // hkv, 0)
        s = (q @ k.transpose(-1, -2)) * D ** -0.5
        s = s.masked_fill(torch.ones(S, S, dtype=torch.bool).triu(1), float("-inf"))
        o = (torch.softmax(s, -1) @ v).transpose(0, 1).reshape(S, hq * D)
        part = F.linear(o, loc[p + "self_attn.o_proj.weight"]) + (h if rank == 0 else 0)     # residual enters once
        dist.all_reduce(part)
        h = part
        y = torch_ref.rmsnorm(h, loc[p + "post_attention_layernorm.weight"], l.eps)
        z = F.silu(F.linear(y, loc[p + "mlp.gate_proj.weight"])) * F.linear(y, loc[p + "mlp.up_proj.weight"])
        part = F.linear(z, loc[p + "mlp.down_proj.weight"]) + (h if rank == 0 else 0)
        dist.all_reduce(part)
        h = part
    hn = torch_ref.rmsnorm(h, loc["model.norm.weight"], l.eps)
    mine = F.linear(hn, loc["lm_head.weight"])
    parts = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(parts, mine)
    got = torch.cat(parts, -1)
    err = ((got - ref).norm() / ref.norm()).item()
    ok = got.shape == ref.shape and err < 1e-4
    # the shards tile the full tensors exactly
    a, b = plan["v_rows"]
    ok = ok and torch.equal(loc["lm_head.weight"], sd["lm_head.weight"][a:b]) and b - a == l.vocab // world
    bad = False
    try:
        shard_plan(ec, 0, 3 if l.heads % 3 else 5)
    except ValueError:
        bad = True
    ok = ok and bad
    dist.barrier()
    dist.destroy_process_group()
    print(f"RANK{rank} {'OK' if ok else 'FAIL'} err={err:.2e}", flush=True)
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()

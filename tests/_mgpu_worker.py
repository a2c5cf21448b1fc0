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
        # and the connector consumes the gathered tokens identically
        a = model.get_model().mm_projector(got[None])
        b = model.get_model().mm_projector(full[None])
        ok &= bool(torch.equal(a, b))
    # ---- the frame-parallel PRODUCT path: encode_images_or_videos / generate with enable_frame_parallel ----------------
    msgs = []
    for name, frames, nvid in (("mid", 6, 1), ("tiny", 4, 2)):
        cfg = synth.CONFIGS[name]
        sd = synth.state_dict(cfg)
        model = build_engine(cfg, sd, dev)
        model.config.num_frames = frames
        g = torch.Generator().manual_seed(7)
        vids = [(torch.randn((frames, 3, cfg.vision.image, cfg.vision.image), generator=g).to(torch.bfloat16).to(dev), "video")
                for _ in range(nvid)]
        want = model.encode_images_or_videos(vids)                      # single-GPU path
        for shard_s1 in (True, False):
            model.enable_frame_parallel(None, shard_s1=shard_s1)
            got = model.encode_images_or_videos(vids)
            same = bool(torch.equal(got, want))
            ok &= same
            if not same:
                msgs.append(f"frame-parallel encode differs ({name}, shard_s1={shard_s1})")
        # generate(): rank 0 decodes, the others return None after the collective
        _, ids = synth.inputs(cfg)
        model.enable_frame_parallel(False)
        ref_tok = model.generate(ids, images=vids[:1], max_new_tokens=3, do_sample=False)
        model.enable_frame_parallel(None, shard_s1=True, llm_rank=0)
        tok = model.generate(ids, images=vids[:1], max_new_tokens=3, do_sample=False)
        if rank == 0:
            ok &= bool(torch.equal(tok, ref_tok))
        else:
            ok &= tok is None
        model.enable_frame_parallel(False)

    # ---- tensor-parallel decoder (model/tp_decoder.py) against the single-GPU engine -------------------------------------
    from helpers import engine_config, rel
    from videollama2_b200.model import VLLMs
    for name in ("tiny", "tiny_qwen2", "mid"):
        cfg = synth.CONFIGS[name]
        if cfg.llm.kv_heads % world or cfg.llm.heads % world:
            continue
        sd = synth.state_dict(cfg)
        ec = engine_config(cfg)
        single = build_engine(cfg, sd, dev)
        tp = VLLMs[ec.model_type].from_state_dict(ec, sd, device=dev, tp_group=True)
        px, ids = synth.inputs(cfg)
        images = [(px.to(dev), "video")]
        a = single(input_ids=ids, attention_mask=torch.ones_like(ids), images=images).logits
        b = tp(input_ids=ids, attention_mask=torch.ones_like(ids), images=images).logits
        err = rel(b, a)
        if not (b.shape == a.shape and err < 1e-2):
            ok = False
            msgs.append(f"TP logits differ ({name}): rel {err:.3e}")
        # the library's own all-reduce kernel (multimem / peer path) against the NCCL path
        for mc in (True, False):
            tp.get_model().decoder.enable_nvls_all_reduce(256, use_multicast=mc)
            for _ in range(2):                              # twice: buffer / epoch reuse
                c = tp(input_ids=ids, attention_mask=torch.ones_like(ids), images=images).logits
            e2 = rel(c, b)
            # the reduced rows are NCCL's bits (peer-load path); the Σx² statistics are summed in another order than
            # vl2_row_sumsq's, and 1-ulp flips of the normalised activations propagate through the layers: ~1e-3
            if not e2 < 3e-3:
                ok = False
                msgs.append(f"own all-reduce kernel differs from the NCCL path ({name}, multicast={mc}): rel {e2:.3e}")
        tp.get_model().decoder._nvls = None
        ta = single.generate(ids, images=images, max_new_tokens=4, do_sample=False)
        tb = tp.generate(ids, images=images, max_new_tokens=4, do_sample=False)
        allb = [torch.empty_like(tb) for _ in range(world)]
        dist.all_gather(allb, tb)
        if not all(torch.equal(t, tb) for t in allb):
            ok = False
            msgs.append(f"TP ranks disagree on the generated ids ({name})")
        if not torch.equal(ta[:, :1], tb[:, :1]):        # first token: same arg-max unless the top-2 gap is inside bf16 noise
            top2 = torch.topk(a[0, -1].float(), 2).values
            if float(top2[0] - top2[1]) > 0.05:
                ok = False
                msgs.append(f"TP first token differs ({name})")
    if msgs:
        print(f"RANK{rank} " + "; ".join(msgs), flush=True)
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    dist.destroy_process_group()
    print(f"RANK{rank} {'OK' if int(flag.item()) == 1 else 'FAIL'}", flush=True)
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()

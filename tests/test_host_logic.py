"""Exact (integer) parity of the host-side token / splice logic with the reference
(videollama2/mm_utils.py:277-302, videollama2/model/videollama2_arch.py:161-263) via the committed fixtures, and live
against the reference functions when /root/reference exists."""
import os

import pytest
import torch

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_tokenizer_multimodal_token_exact():
    from oracle.make_golden import ToyTokenizer
    from oracle import torch_ref
    from videollama2_b200 import mm_utils
    fx = torch.load(os.path.join(GOLD, "tokenizer_multimodal_token.pt"))["tokenizer"]
    tok = ToyTokenizer()
    assert len(fx) >= 6
    for prompt, tag, ref_ids in fx:
        assert mm_utils.tokenizer_multimodal_token(prompt, tok, tag) == ref_ids
        assert torch_ref.tokenizer_multimodal_token(prompt, tok, tag) == ref_ids
        t = mm_utils.tokenizer_multimodal_token(prompt, tok, tag, return_tensors="pt")
        assert t.dtype == torch.long and t.tolist() == ref_ids
    with pytest.raises(ValueError):
        mm_utils.tokenizer_multimodal_token("x", tok, "<video>", return_tensors="np")


@pytest.mark.parametrize("name", ["tiny", "tiny_v35"])
def test_batch_splice_plan_exact(name):
    """Ragged batch of two (placeholder at different positions): new lengths, mask, labels, text-row placement."""
    from oracle import synth
    from videollama2_b200 import mm_utils
    g = torch.load(os.path.join(GOLD, f"{name}.pt"))["splice_batch"]
    cfg = synth.CONFIGS[name]
    ids2 = g["ids"]
    L = cfg.vis_tokens
    plan = mm_utils.build_splice(ids2, [L, L])
    assert plan["max_len"] == g["embeds"].shape[1] and plan["new_len"] == [cfg.prompt - 1 + L] * 2
    mask = mm_utils.spliced_attention_mask(torch.ones_like(ids2, dtype=torch.bool), ids2.shape[1], plan["new_len"], plan["max_len"])
    assert torch.equal(mask, g["mask"])
    labels = mm_utils.spliced_labels(ids2.clone(), ids2, [L, L], plan["max_len"])
    assert torch.equal(labels, g["labels"])
    # text rows land exactly where the reference put them
    table = synth.make_tensor("model.embed_tokens.weight", (cfg.llm.vocab, cfg.llm.hidden), "emb").float()
    emb = g["embeds"].reshape(-1, cfg.llm.hidden)
    for b, src, dst in zip(plan["text_b"], plan["text_src"], plan["text_dst"]):
        assert torch.equal(emb[dst], table[ids2[b, src]])
    assert [(m, b, p, n) for m, b, p, n in plan["mm_dst"]] == [(0, 0, 4, L), (1, 1, 9, L)]


def test_splice_plan_edge_cases():
    from videollama2_b200 import mm_utils
    segs, total, used = mm_utils.splice_plan([5, 6, 7], [10])                 # no placeholder: consumes one slot
    assert segs == [("text", 0, 3)] and total == 3 and used == 1
    segs, total, used = mm_utils.splice_plan([-201], [10])                    # only a placeholder
    assert segs == [("mm", 0, 10)] and total == 10 and used == 1
    segs, total, used = mm_utils.splice_plan([-200, 4, -202], [3, 2])         # image first, audio last
    assert segs == [("mm", 0, 3), ("text", 1, 1), ("mm", 1, 2)] and total == 6 and used == 2
    plan = mm_utils.build_splice(torch.tensor([[1, -201, 2], [3, 4, 5]]), [7, 7])
    assert plan["new_len"] == [9, 3] and plan["max_len"] == 9                 # ragged -> right padding
    m = mm_utils.spliced_attention_mask(torch.ones(2, 3, dtype=torch.bool), 3, plan["new_len"], 9)
    assert m[0].all() and m[1].tolist() == [True] * 3 + [False] * 6


def test_live_reference_splice_and_tokenizer():
    from oracle import ref_loader
    if not ref_loader.available():
        pytest.skip("/root/reference not present")
    import importlib
    ref_loader.load()
    ref_mm = importlib.import_module("videollama2.mm_utils")
    from oracle.make_golden import PROMPTS, ToyTokenizer
    from videollama2_b200 import mm_utils
    tok = ToyTokenizer()
    for p, t in PROMPTS + [("a <video> b <video> c", "<video>"), ("", "<video>")]:
        assert mm_utils.tokenizer_multimodal_token(p, tok, t) == ref_mm.tokenizer_multimodal_token(p, tok, t)


def test_keywords_stopping_criteria():
    from oracle.make_golden import ToyTokenizer
    from videollama2_b200.mm_utils import KeywordsStoppingCriteria
    tok = ToyTokenizer()
    kw = tok("stop now").input_ids[1:]
    sc = KeywordsStoppingCriteria(["stop now"], tok, torch.zeros(1, 3, dtype=torch.long))
    assert sc(torch.tensor([[9, 9] + kw]), None) and not sc(torch.tensor([[9, 9, 9]]), None)


def test_vision_cache_content_key():
    """The vision-feature cache key is a content checksum: equal for a copy, different after a one-element change, a
    different modality or a different shape (videollama2_arch.enable_vision_cache)."""
    from videollama2_b200.model.videollama2_arch import Videollama2MetaForCausalLM as M
    g = torch.Generator().manual_seed(5)
    x = torch.randn((4, 3, 14, 14), generator=g).to(torch.bfloat16)
    k0 = M._content_key([(x, "video")])
    assert M._content_key([(x.clone(), "video")]) == k0
    y = x.clone()
    y[3, 2, 13, 13] = y[3, 2, 13, 13] + 0.5
    assert M._content_key([(y, "video")]) != k0
    assert M._content_key([(x, "image")]) != k0
    assert M._content_key([(x.view(4, 3, 7, 28), "video")]) != k0
    z = x.clone()
    z[0], z[1] = x[1], x[0]                      # a permutation of frames keeps the plain sum, not the weighted one
    assert M._content_key([(z, "video")]) != k0
    odd = torch.arange(5, dtype=torch.uint8)      # byte count not a multiple of 8
    assert M._content_key([(odd, "video")]) == M._content_key([(odd.clone(), "video")])

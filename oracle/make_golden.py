"""Regenerate tests/golden/*.pt by running the REAL reference classes (imported from /root/reference through
oracle/ref_loader.py) on CPU with the deterministic synthetic weights/inputs of oracle/synth.py.
TEST INFRASTRUCTURE.  Run in the build container:  python -m oracle.make_golden
The GPU box has no /root/reference; it consumes the committed fixtures."""
from __future__ import annotations

import os

import torch

from . import ref_loader, synth

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


class ToyTokenizer:
    """Deterministic whitespace tokenizer with the two members tokenizer_multimodal_token uses."""
    bos_token_id = 1

    class _Enc:
        def __init__(self, ids):
            self.input_ids = ids

    def __call__(self, text, add_special_tokens=True):
        ids = [3 + (sum(ord(c) * (i + 1) for i, c in enumerate(w)) % 500) for w in text.split()]
        if add_special_tokens:
            ids = [self.bos_token_id] + ids
        return self._Enc(ids)


PROMPTS = [
    ("<video>\nDescribe the video in detail.", "<video>"),
    ("[INST] <<SYS>> be brief <</SYS>> <video>\nWhat happens? [/INST]", "<video>"),
    ("no visual token at all", "<video>"),
    ("<image> first <image> second", "<image>"),
    ("tail tag <audio>", "<audio>"),
    ("plain text prompt", "not-a-modal-tag"),
]


def main():
    os.makedirs(OUT, exist_ok=True)
    model_mod = ref_loader.load()
    import importlib
    mm_utils = importlib.import_module("videollama2.mm_utils")
    tok = ToyTokenizer()
    fixtures = {"tokenizer": [(p, t, mm_utils.tokenizer_multimodal_token(p, tok, t)) for p, t in PROMPTS]}
    torch.save(fixtures, os.path.join(OUT, "tokenizer_multimodal_token.pt"))

    import sys
    names = sys.argv[1:] or ("tiny", "tiny_qwen2", "tiny_v35", "tiny_siglip")
    for name in names:
        cfg = synth.CONFIGS[name]
        sd = synth.state_dict(cfg)
        px, ids = synth.inputs(cfg)
        out = {"config": name}
        for tag, dt in (("g32", torch.float32), ("hbf16", torch.bfloat16)):
            m = ref_loader.build_reference_model(cfg, dt, sd)
            with torch.no_grad():
                images = [(px.to(dt), "video")]
                feats = m.get_model().get_vision_tower()(px.to(dt))
                mm = m.encode_images_or_videos(images)
                res = m(input_ids=ids, attention_mask=torch.ones_like(ids), images=images)
                out[tag] = {"vit": feats.float(), "mm": mm[0].float(), "logits": res.logits[0].float()}
                if tag == "g32":
                    # index-exact splice fixtures: ragged batch of 2, labels and mask (videollama2_arch.py:161-263)
                    ids2 = torch.stack([ids[0], ids[0].clone()])
                    ids2[1, 4] = 7            # sample 1 has NO placeholder
                    ids2[1, 9] = -201         # ... but one further back
                    mask2 = torch.ones_like(ids2, dtype=torch.bool)
                    lab2 = ids2.clone()
                    r = m.prepare_inputs_labels_for_multimodal(ids2, mask2, None, lab2, [(px.float(), "video"), (px.float(), "video")])
                    out["splice_batch"] = {"ids": ids2, "mask": r[1], "labels": r[4], "embeds": r[3].float()}
                    gen = m.generate(ids, images=images, attention_mask=torch.ones_like(ids), max_new_tokens=4,
                                     do_sample=False, use_cache=True, pad_token_id=0)
                    out["generate_greedy"] = gen
        torch.save(out, os.path.join(OUT, f"{name}.pt"))
        g, h = out["g32"], out["hbf16"]
        rel = lambda a, b: ((a - b).norm() / b.norm()).item()
        print(name, "hbf16-vs-g32:", {k: round(rel(h[k], g[k]), 4) for k in g}, "gen", out["generate_greedy"].tolist())


if __name__ == "__main__":
    main()

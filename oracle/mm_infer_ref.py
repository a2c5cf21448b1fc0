"""Run the REAL reference `mm_infer` (videollama2/__init__.py:39-114) on CPU with a toy tokenizer and a recording fake
model, to pin the prompt / input_ids / generate() arguments that videollama2_b200.mm_infer must reproduce.
TEST INFRASTRUCTURE (build container only: needs /root/reference).  `Tensor.cuda` is patched to the identity for the
duration of the call - the reference moves its tensors to the GPU unconditionally."""
from __future__ import annotations

import importlib.util
import os
import sys

import torch

from . import ref_loader


class ToyChatTokenizer:
    """Deterministic stand-in with the members mm_infer and its helpers touch."""
    bos_token_id = 1
    eos_token = "</s>"
    eos_token_id = 2
    pad_token_id = 0

    class _Enc:
        def __init__(self, ids):
            self.input_ids = ids

    def __call__(self, text, add_special_tokens=True):
        ids = [3 + (sum(ord(c) * (i + 1) for i, c in enumerate(w)) % 5000) for w in text.split()]
        return self._Enc(([self.bos_token_id] if add_special_tokens else []) + ids)

    def apply_chat_template(self, message, tokenize=False, add_generation_prompt=True):
        assert tokenize is False
        s = "".join(f"[{m['role']}] {m['content']} [/{m['role']}] " for m in message)
        return s + ("[assistant]" if add_generation_prompt else "")

    def batch_decode(self, ids, skip_special_tokens=True):
        return [" ".join(str(int(t)) for t in row if not (skip_special_tokens and int(t) in (0, 1, 2))) for row in ids]


class RecordingModel:
    def __init__(self, model_type: str):
        self.config = type("Cfg", (), {"model_type": model_type})()
        self.device = torch.device("cpu")
        self.calls = []

    def generate(self, input_ids, **kw):
        self.calls.append({"input_ids": input_ids.clone(), **kw})
        return torch.tensor([[11, 12, 2, 13]])


def reference_mm_infer():
    ref_loader.load()
    path = os.path.join(ref_loader.REF_ROOT, "videollama2", "__init__.py")
    spec = importlib.util.spec_from_file_location("videollama2._entry", path)
    mod = importlib.util.module_from_spec(spec)
    mod.__package__ = "videollama2"
    sys.modules["videollama2._entry"] = mod
    spec.loader.exec_module(mod)
    return mod.mm_infer


def run_reference(instruct, modal: str, model_type: str, frames=None, **kwargs):
    fn = reference_mm_infer()
    model, tok = RecordingModel(model_type), ToyChatTokenizer()
    orig = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self
    try:
        out = fn(frames, instruct, model, tok, modal=modal, **kwargs)
    finally:
        torch.Tensor.cuda = orig
    return out, model.calls[0]


def summarise(call) -> dict:
    """What must match: ids, mask, the modality tag of `images`, and every scalar generate() argument."""
    d = {"input_ids": call["input_ids"], "attention_mask": call["attention_mask"],
         "images_modal": None if call["images"] is None else call["images"][0][1],
         "n_stopping": len(call["stopping_criteria"])}
    for k in ("do_sample", "temperature", "max_new_tokens", "top_p", "use_cache", "pad_token_id"):
        d[k] = call[k]
    return d


CASES = [
    ("Describe the video in detail.", "video", "videollama2_mistral", {}),
    ("What is in the image?", "image", "videollama2_qwen2", {"max_new_tokens": 64}),
    ("Just text, no pixels.", "text", "videollama2", {"do_sample": True}),
    ([{"role": "user", "content": "first turn"}, {"role": "assistant", "content": "ok"},
      {"role": "user", "content": "second turn"}], "video", "videollama2_qwen2", {"top_p": 0.5, "temperature": 0.7}),
]


def main():
    out = []
    for instruct, modal, mtype, kw in CASES:
        frames = None if modal == "text" else torch.zeros((2, 3, 4, 4))
        text, call = run_reference(instruct, modal, mtype, frames, **kw)
        out.append({"instruct": instruct, "modal": modal, "model_type": mtype, "kwargs": kw, "text": text,
                    "call": summarise(call)})
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "mm_infer.pt")
    torch.save(out, path)
    print("wrote", path, [o["text"] for o in out])


if __name__ == "__main__":
    main()

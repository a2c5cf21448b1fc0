"""The BASELINE configurations at FULL depth (23 ViT layers + 8 RegStage blocks + 32 / 28 decoder layers, S = 876 / 1776)
against goldens computed by the real reference classes on CPU (oracle/make_golden_full.py -> tests/golden/full_cfg*.pt).

The 7B synthetic checkpoint is regenerated here from its seed on the host RNG (presets.synthetic_state_dict: the same
bytes the goldens were computed with), the engine runs pixels + ids -> logits with its own kernels, and the taps (ViT
output, connector output, decoder layers {first, middle, last}, last-position logits) are compared at the fixture's rows:
relL2 <= max(1e-2, 1.25 x the reference's own bf16-vs-fp32 error there), same arg-max token.  One engine build per
decoder family (cfg1 and cfg2 share the Mistral-7B weights)."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu


def _need(name):
    from videollama2_b200 import selfcheck
    p = selfcheck.fixture_path(name)
    if not os.path.exists(p):
        pytest.skip(f"{p} not generated (python -m oracle.make_golden_full)")


def _engine(llm, cuda, frames):
    from videollama2_b200 import presets
    from videollama2_b200.model import VLLMs
    cfg = presets.make_config(llm, frames)
    sd = presets.synthetic_state_dict(cfg, cuda, threads=min(16, os.cpu_count() or 1))
    model = VLLMs[cfg.model_type].from_state_dict(cfg, sd, device=cuda)
    del sd
    torch.cuda.empty_cache()
    return cfg, model


def _report(res):
    return {k: (round(v["rel_l2"], 5), round(v["bar"], 5)) for k, v in res["taps"].items()}, res["argmax"]


@pytest.fixture(scope="module")
def mistral7b(cuda):
    from videollama2_b200 import presets
    _need("cfg2")
    cfg, model = _engine(presets.MISTRAL_7B, cuda, 16)
    yield cfg, model
    del model
    torch.cuda.empty_cache()


@pytest.mark.parametrize("name,frames,prompt", [("cfg2", 16, 256), ("cfg1", 8, 32)])
def test_mistral7b_full_depth(mistral7b, name, frames, prompt):
    from videollama2_b200 import presets, selfcheck
    _need(name)
    cfg, model = mistral7b
    model.config.num_frames = frames
    px, ids = presets.synthetic_inputs(cfg, frames, prompt)
    res = selfcheck.fulldepth_check(model, name, px, ids)
    print(name, _report(res))
    assert res["ok"], res


def test_mistral7b_full_depth_graphs_and_generate(mistral7b):
    """The graph-replayed stages (what bench.py times) give the SAME last-position token as the eager tapped pass, and
    generate() returns it."""
    from videollama2_b200 import presets
    cfg, model = mistral7b
    model.config.num_frames = 16
    px, ids = presets.synthetic_inputs(cfg, 16, 256)
    fx = torch.load(os.path.join(os.path.dirname(__file__), "golden", "full_cfg2.pt"), map_location="cpu")
    dev = model.device
    eager = model.generate(ids, images=[(px.to(dev), "video")], max_new_tokens=1, do_sample=False)
    model.enable_cuda_graphs(True)
    try:
        graphed = model.generate(ids, images=[(px.to(dev), "video")], max_new_tokens=1, do_sample=False)
    finally:
        model.enable_cuda_graphs(False)
    assert torch.equal(eager, graphed)
    gap = float(fx["g32"]["logits_last"][fx["argmax_g32"]] - fx["g32"]["logits_last"][int(eager[0, 0])])
    assert gap <= 2.0 * float(fx["logit_noise_absmax"])


def test_qwen2_7b_full_depth(cuda):
    """Config 3: Qwen2-7B at its real geometry (28 q / 4 kv heads = GQA group 7, H = 3584, I = 18944, V = 152064, q/k/v
    bias) behind the CLIP tower + STC connector."""
    from videollama2_b200 import presets, selfcheck
    _need("cfg3")
    cfg, model = _engine(presets.QWEN2_7B, cuda, 16)
    px, ids = presets.synthetic_inputs(cfg, 16, 256)
    res = selfcheck.fulldepth_check(model, "cfg3", px, ids)
    print("cfg3", _report(res))
    del model
    torch.cuda.empty_cache()
    assert res["ok"], res

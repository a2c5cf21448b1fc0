"""Videollama2Qwen2ForCausalLM (videollama2/model/videollama2_qwen2.py:45-151): identical wrapper over the Qwen2
decoder (q/k/v bias, HF:qwen2/modeling_qwen2.py:200-202); the engine picks the biases up from the state dict."""
from .config import Videollama2Config
from .videollama2_mistral import Videollama2MistralForCausalLM


class Videollama2Qwen2Config(Videollama2Config):
    model_type = "videollama2_qwen2"


class Videollama2Qwen2ForCausalLM(Videollama2MistralForCausalLM):
    config_class = Videollama2Qwen2Config

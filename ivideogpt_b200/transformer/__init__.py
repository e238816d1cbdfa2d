from .action_model import HeadModelWithAction
from .llama_b200 import B200LlamaForCausalLM, register

register()  # AutoModelForCausalLM -> B200LlamaForCausalLM for LlamaConfig (the drop-in seam)

__all__ = ["HeadModelWithAction", "B200LlamaForCausalLM"]

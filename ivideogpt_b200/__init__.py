"""ivideogpt_b200 -- B200-native (sm_100a) implementation of iVideoGPT's next-frame-prediction hot path.

    from ivideogpt_b200.vq_model import CompressiveVQModel          # ctx_vqgan tokenizer (encode / VQ / decode)
    from ivideogpt_b200.transformer import HeadModelWithAction       # + registers the Llama class with transformers

Everything that computes is a hand-written CUDA kernel in csrc/, reached through the C ABI declared in
include/ivgpt_b200.h (ctypes binding in _lib.py).  There is no CPU path and no PyTorch fallback.
"""
__version__ = "0.1.0"

from .tokenizer import CompressiveVQModel

__all__ = ["CompressiveVQModel"]

"""Drop-in alias: `ivideogpt.vq_model` / `ivideogpt.transformer` import paths of thuml/iVideoGPT
(ivideogpt/vq_model/__init__.py:1-3, ivideogpt/transformer/__init__.py:1) served by ivideogpt_b200."""
import importlib

from ivideogpt_b200.vq_model import CompressiveVQModel  # noqa: F401

_REFERENCE_OWN = {"Discriminator": "discriminator", "LPIPS": "lpips"}


def __getattr__(name):
    # Discriminator / LPIPS (tokenizer GAN training, reference __init__.py:2-3) are outside the B200 hot path.  When this file
    # replaces the reference's ivideogpt/vq_model/__init__.py the reference's own discriminator.py / lpips.py
    # sit next to it and keep being served from here; in this repository alone they do not exist.
    if name in _REFERENCE_OWN:
        try:
            return getattr(importlib.import_module("." + _REFERENCE_OWN[name], __name__), name)
        except ModuleNotFoundError as e:
            if e.name is None or not e.name.endswith(_REFERENCE_OWN[name]):
                raise           # the reference module exists but one of ITS dependencies is missing: say so
            raise ImportError(f"ivideogpt.vq_model.{name} belongs to tokenizer GAN training, which is outside the B200 "
                              "hot path (SURVEY.md section 2, rows 12-13); use the reference implementation for it.") from None
    raise AttributeError(name)

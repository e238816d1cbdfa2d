"""Drop-in alias: `ivideogpt.vq_model` / `ivideogpt.transformer` import paths of thuml/iVideoGPT
(ivideogpt/vq_model/__init__.py:1-3, ivideogpt/transformer/__init__.py:1) served by ivideogpt_b200."""

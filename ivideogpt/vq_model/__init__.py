from ivideogpt_b200.vq_model import CompressiveVQModel  # noqa: F401


def __getattr__(name):
    if name in ("Discriminator", "LPIPS"):
        raise ImportError(f"ivideogpt.vq_model.{name} belongs to tokenizer GAN training, which is outside the B200 "
                          "hot path (SURVEY.md section 2, rows 12-13); use the reference implementation for it.")
    raise AttributeError(name)

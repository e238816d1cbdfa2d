from ivideogpt_b200.transformer import HeadModelWithAction  # noqa: F401

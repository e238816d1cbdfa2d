"""ModelMixin: an nn.Module (loading / saving is not exercised by the golden generator)."""
import torch


class ModelMixin(torch.nn.Module):
    pass

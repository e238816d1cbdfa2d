from collections import OrderedDict

import torch


class BaseOutput(OrderedDict):
    """Dataclass-style output container (attribute access only is exercised here)."""

    def __post_init__(self):
        pass


def is_torch_version(op, version):
    return True

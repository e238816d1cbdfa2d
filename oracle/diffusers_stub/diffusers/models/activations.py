import torch.nn as nn


def get_activation(name: str):
    table = {"silu": nn.SiLU, "swish": nn.SiLU, "relu": nn.ReLU, "gelu": nn.GELU, "mish": nn.Mish}
    return table[name.lower()]()

"""unicore.utils.get_activation_fn restated (used at reference models/mm_model.py:66,115)."""
import torch
import torch.nn.functional as F


def get_activation_fn(name):
    table = {
        "relu": F.relu,
        "gelu": F.gelu,            # exact erf form
        "tanh": torch.tanh,
        "linear": lambda x: x,
    }
    if name not in table:
        raise RuntimeError("--activation-fn {} not supported".format(name))
    return table[name]

"""unicore.models.BaseUnicoreModel stand-in (imported, never used, at models/mm_model.py:15)."""
import torch.nn as nn


class BaseUnicoreModel(nn.Module):
    pass

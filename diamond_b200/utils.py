"""The two helpers of the reference's src/utils.py that the mirrored models need (init_lstm utils.py:184-196)."""
from typing import Any, Dict, Tuple

import torch.nn as nn
from torch import Tensor

LossAndLogs = Tuple[Tensor, Dict[str, Any]]


def init_lstm(model: nn.Module) -> None:
    for name, p in model.named_parameters():
        if "weight_ih" in name:
            nn.init.xavier_uniform_(p.data)
        elif "weight_hh" in name:
            nn.init.orthogonal_(p.data)
        elif "bias_ih" in name:
            p.data.fill_(0)
            n = p.size(0)
            p.data[(n // 4):(n // 2)].fill_(1)  # forget-gate bias
        elif "bias_hh" in name:
            p.data.fill_(0)

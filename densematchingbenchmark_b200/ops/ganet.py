"""GANet aggregation layers as nn.Modules (no reference code exists in the dmb snapshot --
SURVEY.md section 0.1; semantics are fixed by oracle/dmb_oracle.py:sga / lga)."""
import torch.nn as nn

from . import functional as F_
from .autograd import forbid_grad


class SGA(nn.Module):
    """Semi-global aggregation: forward(x [B,C,D,H,W], guidance [B,4*5*C,H,W]) -> [B,C,D,H,W]."""

    def forward(self, x, guidance):
        forbid_grad("SGA", x, guidance)
        return F_.sga(x, guidance)


class LGA(nn.Module):
    """Local guided aggregation: forward(x [B,D,H,W], guidance [B,3*(2r+1)^2,H,W]) -> [B,D,H,W]."""

    def __init__(self, radius=2):
        super(LGA, self).__init__()
        self.radius = radius

    def forward(self, x, guidance):
        forbid_grad("LGA", x, guidance)
        return F_.lga(x, guidance, self.radius)

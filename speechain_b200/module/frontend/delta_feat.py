"""Delta / delta-delta features.  API mirror of speechain/module/frontend/delta_feat.py."""
import torch

from ...host import ops
from ..abs import Module
from ._common import ConstHolder


class DeltaFeature(Module):
    def module_init(self, delta_order: int = 1, delta_N: int = 2):
        assert isinstance(delta_order, int) and delta_order in [1, 2], \
            f"delta_order must be an integer of 1 or 2, but got delta_order={delta_order}"
        assert isinstance(delta_N, int) and delta_N >= 1, "delta_N must be a positive integer equal to or larger than 1"
        self.delta_order, self.delta_N = delta_order, delta_N
        w = torch.arange(-delta_N, delta_N + 1) / (2 * sum(i ** 2 for i in range(1, delta_N + 1)))
        self.delta_filters = ConstHolder(w.reshape(1, 1, -1, 1))  # same key/shape as delta_feat.py:40-54

    def forward(self, feat: torch.Tensor, feat_len: torch.Tensor):
        return ops.delta(feat, self.delta_order, self.delta_N), feat_len

    def __repr__(self) -> str:
        return f"{self.__class__.__name__}(\ndelta_order={self.delta_order}, delta_N={self.delta_N}\n)"

import torch.nn as tnn

from ...utils import scatter


class _Reduce(tnn.Module):
    reduce = "sum"

    def forward(self, x, index=None, ptr=None, dim_size=None, dim=-2):
        return scatter(x, index, 0, dim_size, self.reduce)


class SumAggregation(_Reduce):
    reduce = "sum"


class MaxAggregation(_Reduce):
    reduce = "max"

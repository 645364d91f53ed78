"""GraphNorm (Cai et al. 2020) with PyG's parameterisation, restated.

forward(x, batch):  mean = segment_mean(x, batch)
                    out  = x - mean[batch] * mean_scale
                    var  = segment_mean(out * out, batch)
                    y    = weight * out / sqrt(var + eps)[batch] + bias
Parameters: weight (ones), bias (zeros), mean_scale (ones); eps = 1e-5.
"""
import torch
import torch.nn as tnn

from ...utils import scatter


class GraphNorm(tnn.Module):
    def __init__(self, in_channels, eps=1e-5):
        super().__init__()
        self.in_channels = in_channels
        self.eps = eps
        self.weight = tnn.Parameter(torch.ones(in_channels))
        self.bias = tnn.Parameter(torch.zeros(in_channels))
        self.mean_scale = tnn.Parameter(torch.ones(in_channels))

    def forward(self, x, batch=None, batch_size=None):
        if batch is None:
            batch = x.new_zeros(x.size(0), dtype=torch.long)
        if batch_size is None:
            batch_size = int(batch.max()) + 1
        mean = scatter(x, batch, 0, batch_size, "mean")
        out = x - mean.index_select(0, batch) * self.mean_scale
        var = scatter(out.pow(2), batch, 0, batch_size, "mean")
        std = (var + self.eps).sqrt().index_select(0, batch)
        return self.weight * out / std + self.bias

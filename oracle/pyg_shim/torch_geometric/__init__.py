"""Minimal `torch_geometric` stand-in.  TEST INFRASTRUCTURE ONLY.

The reference (hhcaz/CNS) does all hot-path arithmetic through PyTorch Geometric,
which is un-vendored, un-pinned (absent from the reference's requirements.txt;
pickle metadata in checkpoints/cns.pth points at PyG ~2.1-2.2) and not installable
in this image.  This package restates the published semantics of exactly the PyG
symbols the reference touches so that `/root/reference/cns/models/graph_vs.py` and
`/root/reference/cns/midend/graph_gen.py` import UNMODIFIED:

  torch_geometric.data.Data / Batch              (graph_gen.py:3,68 ; dataset.py:5,208)
  torch_geometric.nn.MessagePassing              (graph_vs.py:8,46)
  torch_geometric.nn.PointTransformerConv        (graph_vs.py:8,195)
  torch_geometric.nn.AttentionalAggregation      (graph_vs.py:8,259)
  torch_geometric.nn.norm.GraphNorm              (graph_vs.py:7,132)

Nothing in `cns_b200/` imports this.  Only `oracle/` and `tests/` do.
"""
__version__ = "0.0-shim"

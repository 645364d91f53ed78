from .. import PointTransformerConv  # noqa: F401

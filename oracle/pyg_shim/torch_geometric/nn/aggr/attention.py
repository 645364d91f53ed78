from .. import AttentionalAggregation  # noqa: F401

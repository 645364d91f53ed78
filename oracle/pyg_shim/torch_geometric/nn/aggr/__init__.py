"""Import paths of PyG's aggregation modules (pickled inside MessagePassing modules).  TEST INFRASTRUCTURE ONLY."""
from .basic import MaxAggregation, SumAggregation  # noqa: F401
